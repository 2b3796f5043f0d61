from skeletor_b200.time_steppers.horowitz import TimeStepper  # noqa: F401
