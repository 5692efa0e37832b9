from rl.parameters.parameters import (DynamicParameter, ConstantParameter, ScheduleWrapper, ExponentialDecay, StepDecay,
                                      PolynomialDecay, LearningRateSchedule)
