"""Dynamic step-dependent parameters: same interface as the reference's rl/parameters/parameters.py:9-92.

The reference wraps `tf.keras.optimizers.schedules`; TensorFlow is not a dependency here, so the three
schedules it uses are restated (Keras formulas) behind the same class names and constructor arguments."""
import math
from typing import Union


class LearningRateSchedule:
    """Stand-in for tf.keras.optimizers.schedules.LearningRateSchedule: a callable of the step."""

    def __call__(self, step):
        raise NotImplementedError

    def get_config(self) -> dict:
        return {}


class _ExponentialDecay(LearningRateSchedule):
    def __init__(self, initial_learning_rate, decay_steps, decay_rate, staircase=False):
        self.initial_learning_rate, self.decay_steps, self.decay_rate, self.staircase = \
            initial_learning_rate, decay_steps, decay_rate, staircase

    def __call__(self, step):
        p = step / self.decay_steps
        if self.staircase:
            p = math.floor(p)
        return self.initial_learning_rate * (self.decay_rate ** p)

    def get_config(self):
        return dict(initial_learning_rate=self.initial_learning_rate, decay_steps=self.decay_steps,
                    decay_rate=self.decay_rate, staircase=self.staircase)


class _PolynomialDecay(LearningRateSchedule):
    def __init__(self, initial_learning_rate, decay_steps, end_learning_rate=0.0001, power=1.0, cycle=False):
        self.initial_learning_rate, self.decay_steps, self.end_learning_rate, self.power, self.cycle = \
            initial_learning_rate, decay_steps, end_learning_rate, power, cycle

    def __call__(self, step):
        decay_steps = self.decay_steps
        if self.cycle:
            decay_steps = decay_steps * max(1.0, math.ceil(step / decay_steps))
        else:
            step = min(step, decay_steps)
        p = step / decay_steps
        return (self.initial_learning_rate - self.end_learning_rate) * ((1 - p) ** self.power) + self.end_learning_rate

    def get_config(self):
        return dict(initial_learning_rate=self.initial_learning_rate, decay_steps=self.decay_steps,
                    end_learning_rate=self.end_learning_rate, power=self.power, cycle=self.cycle)


class DynamicParameter:
    """Interface for learning rate schedule wrappers as dynamic-parameters (parameters.py:9-41)."""

    def __init__(self):
        self.value = 0
        self.step = 0

    @staticmethod
    def create(value: Union[float, LearningRateSchedule, 'DynamicParameter'], **kwargs):
        if isinstance(value, float):
            return ConstantParameter(value)
        if isinstance(value, DynamicParameter):
            return value
        if isinstance(value, LearningRateSchedule):
            return ScheduleWrapper(schedule=value, **kwargs)
        assert isinstance(value, DynamicParameter) or isinstance(value, ScheduleWrapper)
        return value

    def __call__(self, *args, **kwargs):
        return self.value

    def serialize(self) -> dict:
        return dict(step=int(self.step))

    def on_episode(self):
        self.step += 1

    def load(self, config: dict):
        self.step = config.get('step', 0)

    def get_config(self) -> dict:
        return {}


class ScheduleWrapper(LearningRateSchedule, DynamicParameter):
    """A wrapper for learning rate schedules (parameters.py:45-58)."""

    def __init__(self, schedule: LearningRateSchedule, min_value=1e-4):
        DynamicParameter.__init__(self)
        self.schedule = schedule
        self.min_value = min_value

    def __call__(self, *args, **kwargs):
        self.value = max(self.min_value, self.schedule(self.step))
        return self.value

    def get_config(self) -> dict:
        return self.schedule.get_config()


class ConstantParameter(DynamicParameter):
    def __init__(self, value: float):
        super().__init__()
        self.value = value

    def __call__(self, *args, **kwargs):
        return self.value

    def serialize(self) -> dict:
        return {}


class ExponentialDecay(ScheduleWrapper):
    def __init__(self, initial_value: float, decay_steps: int, decay_rate: float, staircase=False, min_value=0.0):
        super().__init__(schedule=_ExponentialDecay(initial_value, decay_steps, decay_rate, staircase), min_value=min_value)


class StepDecay(ScheduleWrapper):
    def __init__(self, initial_value: float, decay_steps: int, decay_rate: float, min_value=1e-4):
        super().__init__(schedule=_ExponentialDecay(initial_value, decay_steps, decay_rate, staircase=True), min_value=min_value)


class PolynomialDecay(ScheduleWrapper):
    def __init__(self, initial_value: float, end_value: float, decay_steps: int, power=1.0, cycle=False):
        super().__init__(schedule=_PolynomialDecay(initial_value, decay_steps, end_value, power, cycle))
