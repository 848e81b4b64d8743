"""jogramop_framework_b200 -- B200-native batched collision-query engine for the jogramop scenarios.

Drop-in surface (same names as the reference's modules):
    from jogramop_framework_b200.scenario import Scenario
    from jogramop_framework_b200.simulation import FrankaRobot, GraspingSimulator
Engine: jogramop_framework_b200.engine.CollisionEngine over the C-ABI of include/jogramop_b200.h.
"""
__version__ = '0.1.0'
