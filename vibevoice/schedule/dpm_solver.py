from vibevoice_b200.schedule import DPMSolverMultistepScheduler  # noqa: F401
