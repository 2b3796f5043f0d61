"""Run-directory bookkeeping, `.npz` field snapshots and the run log.

Same files and keys as the reference's `skeletor.io.IO` (io.py:1-147): `info.p` (pickled
dict of the caller's int/float variables, git commit, hostname, start time, rank count,
run time), `fields.NNNN.npz` (rho, Ex, Ey, x, y, t — global arrays gathered to rank 0)
and `skeletor.log`.  Diagnostics path: the fields are copied to the host and gathered
with the communicator's object allgather.  The reference's VizSchema HDF5 writer/reader
(writer.py, reader.py, merge.py) needs h5py, which this image does not have.
"""
import os
import pickle
import shutil
import socket
import subprocess
import time
from datetime import datetime

import numpy as np

_STAMP = '%d/%m/%Y at %H:%M:%S'
_LOG = 'skeletor.log'


def _git_commit():
    try:
        out = subprocess.check_output(['git', 'rev-parse', 'HEAD'], stderr=subprocess.DEVNULL)
        return out.strip().decode('utf-8')
    except Exception:
        return None


class IO:
    def __init__(self, data_folder, local_vars, experiment, tag='', comm=None):
        if comm is None:
            from .comm import COMM_WORLD as comm
        self.comm = comm
        self.data_folder = data_folder if data_folder.endswith('/') else data_folder + '/'
        self.snap = 0                       # snapshot counter
        self.wt = time.time()               # for the total run time
        if comm.rank != 0:
            return
        os.makedirs(self.data_folder, exist_ok=True)
        if experiment and os.path.exists(experiment):
            shutil.copy(experiment, self.data_folder + 'experiment.py')
        started = datetime.now().strftime(_STAMP)
        info = dict(experiment=experiment, git_commit=_git_commit(),
                    hostname=socket.gethostname(), simulation_start=started, tag=tag)
        # every plain number in the caller's namespace is a run parameter
        info.update({k: v for k, v in local_vars.items()
                     if type(v) in (float, np.float64, int) and k != 'idproc'})
        info['MPI'] = comm.size
        self._save_info(info)
        with open(_LOG, 'w') as f:
            f.write('Simulation started on ' + started + '\n\n')
            f.write('Contents of info.p is printed below \n')
            f.writelines('{} = {} \n'.format(k, v) for k, v in info.items())
            f.write('\n\nEntering main simulation loop \n')

    def _save_info(self, info):
        with open(self.data_folder + 'info.p', 'wb') as f:
            pickle.dump(info, f)

    def set_outputrate(self, dt):
        self.dt = dt

    def concatenate(self, arr):
        """global array from the slabs of all ranks (available on every rank)"""
        return np.concatenate(self.comm.allgather(np.asarray(arr)))

    def output_fields(self, sources, E, grid, t):
        """write charge density and in-plane electric field as fields.NNNN.npz"""
        rho = self.concatenate(sources.rho.trim())
        Eg = self.concatenate(E.trim())
        if self.comm.rank == 0:
            np.savez('{}fields.{:04d}.npz'.format(self.data_folder, self.snap),
                     rho=rho, Ex=Eg['x'], Ey=Eg['y'], x=grid.x, y=grid.y, t=t)
        self.snap += 1

    def log(self, it, t, dt):
        if self.comm.rank == 0:
            with open(_LOG, 'a') as f:
                f.write('step {0}\ttime {1}\tdt {2}\n'.format(it, t, dt))

    def finished(self):
        """record the run time in info.p and the log, move the log into the run folder"""
        seconds = time.time() - self.wt
        if self.comm.rank != 0:
            return
        with open(self.data_folder + 'info.p', 'rb') as f:
            info = pickle.load(f)
        info['seconds'] = seconds
        self._save_info(info)
        minutes, s = divmod(seconds, 60)
        hours, m = divmod(minutes, 60)
        d, h = divmod(hours, 24)
        with open(_LOG, 'a') as f:
            f.write('Simulation ended on ' + datetime.now().strftime(_STAMP) + '\n')
            f.write('Time elapsed was {} days {} hours {} minutes {} seconds'.format(d, h, m, s))
        shutil.move(_LOG, self.data_folder + _LOG)
