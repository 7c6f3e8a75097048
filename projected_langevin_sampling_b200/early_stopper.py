"""EarlyStopper of the reference's training loops (experiments/early_stopper.py:4-24): host-side logic, no arithmetic."""
import math


class EarlyStopper:
    """Stop on a non-finite loss, or once the loss has not improved on its minimum for `patience` of accumulated
    simulated time (the sum of the step sizes of the non-improving epochs); an improvement resets the clock."""

    def __init__(self, patience: float = 1e-4):
        self.patience = patience
        self.simulation_time = 0
        self.min_loss = float("inf")

    def should_stop(self, loss: float, step_size: float) -> bool:
        if not math.isfinite(loss):
            return True
        if loss >= self.min_loss:
            self.simulation_time += step_size
            return self.simulation_time >= self.patience
        self.min_loss = loss
        self.simulation_time = 0
        return False
