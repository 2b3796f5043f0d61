from skeletor_b200.time_steppers.predictor_corrector import TimeStepper  # noqa: F401
