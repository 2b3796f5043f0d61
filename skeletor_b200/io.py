"""Run-directory bookkeeping and .npz field snapshots (reference skeletor/io.py:1-147).
Diagnostics path: fields are gathered to the host through the communicator's object
allgather, rank 0 writes.  The VizSchema HDF5 writer/reader of the reference
(writer.py, reader.py, merge.py) needs h5py, which this image does not have."""
import os
import pickle
import shutil
import socket
import subprocess
import time
from datetime import datetime

import numpy as np


class IO:
    def __init__(self, data_folder, local_vars, experiment, tag='', comm=None):
        """Creates the output directory, copies the experiment script into it and saves
        a dictionary with the run parameters (io.py:2-76)."""
        if comm is None:
            from .comm import COMM_WORLD as comm
        self.comm = comm
        if data_folder[-1] != '/':
            data_folder += '/'
        if comm.rank == 0:
            os.makedirs(data_folder, exist_ok=True)
            if experiment and os.path.exists(experiment):
                shutil.copy(experiment, data_folder + 'experiment.py')
            info = {'experiment': experiment}
            try:
                git_commit = subprocess.check_output(
                    ['git', 'rev-parse', 'HEAD'], stderr=subprocess.DEVNULL)
                info['git_commit'] = git_commit.strip().decode('utf-8')
            except Exception:
                info['git_commit'] = None
            info['hostname'] = socket.gethostname()
            simulation_start = datetime.now().strftime('%d/%m/%Y at %H:%M:%S')
            info['simulation_start'] = simulation_start
            # tag which can be used to group simulations together
            info['tag'] = tag
            # all int / float variables of the caller's namespace
            for key, val in local_vars.items():
                if type(val) in (float, np.float64, int):
                    info[key] = val
            info['MPI'] = comm.size
            info.pop('idproc', None)
            pickle.dump(info, open(data_folder + 'info.p', 'wb'))
            with open('skeletor.log', 'w') as f:
                f.write('Simulation started on ' + simulation_start + '\n\n')
                f.write('Contents of info.p is printed below \n')
                for key in info:
                    f.write(key + ' = {} \n'.format(info[key]))
                f.write('\n\nEntering main simulation loop \n')
        self.data_folder = data_folder
        self.snap = 0
        self.wt = time.time()

    def set_outputrate(self, dt):
        self.dt = dt

    def concatenate(self, arr):
        """Concatenate local arrays to obtain global arrays on every rank."""
        return np.concatenate(self.comm.allgather(np.asarray(arr)))

    def output_fields(self, sources, E, grid, t):
        """Output charge density and electric field (io.py:91-108)"""
        global_rho = self.concatenate(sources.rho.trim())
        global_E = self.concatenate(E.trim())
        if self.comm.rank == 0:
            np.savez(self.data_folder + 'fields.{:04d}.npz'.format(self.snap),
                     rho=global_rho, Ex=global_E['x'], Ey=global_E['y'],
                     x=grid.x, y=grid.y, t=t)
        self.snap += 1

    def log(self, it, t, dt):
        if self.comm.rank == 0:
            with open('skeletor.log', 'a') as f:
                f.write('step {0}\ttime {1}\tdt {2}\n'.format(it, t, dt))

    def finished(self):
        """Write elapsed time to the log and move it to the data directory"""
        seconds = time.time() - self.wt
        if self.comm.rank == 0:
            info = pickle.load(open(self.data_folder + 'info.p', 'rb'))
            info['seconds'] = seconds
            pickle.dump(info, open(self.data_folder + 'info.p', 'wb'))
            endtime = datetime.now().strftime('%d/%m/%Y at %H:%M:%S')
            m, s = divmod(seconds, 60)
            h, m = divmod(m, 60)
            d, h = divmod(h, 24)
            with open('skeletor.log', 'a') as f:
                f.write('Simulation ended on ' + endtime + '\n')
                msg = 'Time elapsed was {} days {} hours {} minutes {} seconds'
                f.write(msg.format(d, h, m, s))
            shutil.move('skeletor.log', self.data_folder + 'skeletor.log')
