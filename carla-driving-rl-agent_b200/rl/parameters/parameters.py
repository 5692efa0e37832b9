"""Episode-indexed hyper-parameters (learning rates, clip ratio, entropy strength, advantage scale).

Public names, constructor arguments and the serialised form (`{'step': n}` inside config.json) follow the reference's
rl/parameters/parameters.py:9-92 so that `core/learning.py`'s stage definitions run unchanged.  The reference delegates
the decay formulas to `tf.keras.optimizers.schedules`; TensorFlow is not a dependency here, so the two formulas it uses
(exponential / staircase and polynomial, as documented by Keras) are evaluated directly."""
import math


class LearningRateSchedule:
    """A pure function of the step with a describing config -- the role `tf.keras...LearningRateSchedule` plays."""

    def __init__(self, formula=None, **config):
        self._formula, self._config = formula, config

    def __call__(self, step):
        if self._formula is None:
            raise NotImplementedError
        return self._formula(step, **self._config)

    def get_config(self) -> dict:
        return dict(self._config)


def _exponential(step, initial_learning_rate, decay_steps, decay_rate, staircase):
    exponent = step / decay_steps
    return initial_learning_rate * decay_rate ** (math.floor(exponent) if staircase else exponent)


def _polynomial(step, initial_learning_rate, decay_steps, end_learning_rate, power, cycle):
    horizon = decay_steps * max(1.0, math.ceil(step / decay_steps)) if cycle else decay_steps
    remaining = 1.0 - min(step, horizon) / horizon
    return (initial_learning_rate - end_learning_rate) * remaining ** power + end_learning_rate


class DynamicParameter:
    """`value` now, `step` = episodes seen; calling the parameter refreshes and returns `value`."""

    def __init__(self):
        self.value, self.step = 0, 0

    @staticmethod
    def create(value, **kwargs):
        """float -> constant, schedule -> wrapped schedule, parameter -> itself"""
        if isinstance(value, DynamicParameter):
            return value
        if isinstance(value, LearningRateSchedule):
            return ScheduleWrapper(schedule=value, **kwargs)
        if isinstance(value, float):
            return ConstantParameter(value)
        raise AssertionError(f'cannot make a DynamicParameter from {type(value).__name__}')

    def __call__(self, *args, **kwargs):
        return self.value

    def on_episode(self):
        self.step += 1

    def serialize(self) -> dict:
        return {'step': int(self.step)}

    def load(self, config: dict):
        self.step = config.get('step', 0)

    def get_config(self) -> dict:
        return {}


class ConstantParameter(DynamicParameter):
    def __init__(self, value: float):
        super().__init__()
        self.value = value

    def serialize(self) -> dict:          # nothing to restore
        return {}


class ScheduleWrapper(LearningRateSchedule, DynamicParameter):
    """A schedule evaluated at the parameter's own step, floored at `min_value`."""

    def __init__(self, schedule: LearningRateSchedule, min_value=1e-4):
        DynamicParameter.__init__(self)
        self.schedule, self.min_value = schedule, min_value

    def __call__(self, *args, **kwargs):
        self.value = max(self.min_value, self.schedule(self.step))
        return self.value

    def get_config(self) -> dict:
        return self.schedule.get_config()


class ExponentialDecay(ScheduleWrapper):
    def __init__(self, initial_value: float, decay_steps: int, decay_rate: float, staircase=False, min_value=0.0):
        super().__init__(LearningRateSchedule(_exponential, initial_learning_rate=initial_value, decay_steps=decay_steps,
                                              decay_rate=decay_rate, staircase=staircase), min_value=min_value)


class StepDecay(ScheduleWrapper):
    def __init__(self, initial_value: float, decay_steps: int, decay_rate: float, min_value=1e-4):
        super().__init__(LearningRateSchedule(_exponential, initial_learning_rate=initial_value, decay_steps=decay_steps,
                                              decay_rate=decay_rate, staircase=True), min_value=min_value)


class PolynomialDecay(ScheduleWrapper):
    def __init__(self, initial_value: float, end_value: float, decay_steps: int, power=1.0, cycle=False):
        super().__init__(LearningRateSchedule(_polynomial, initial_learning_rate=initial_value, decay_steps=decay_steps,
                                              end_learning_rate=end_value, power=power, cycle=cycle))
