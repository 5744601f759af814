"""pypybox2d_b200 -- B200-native batched `World.step` behind the pypybox2d API.

Drop-in surface (names as `pypybox2d/__init__.py:33-57` re-exports them): `World`, `Body`,
`Fixture`, `Circle`, `Polygon`, `Edge`, `Vec2`, the exception types.  The step itself runs in
hand-written sm_100a CUDA kernels (`csrc/`) reached through the C-ABI in `include/b2b.h`;
there is no CPU implementation of the step in this package.
"""
from . import settings, shapes, scenes, schema
from .common import Vec2, AABB, Transform, PhysicsError, LockedError, EmptyFixtureError, scalar_cross, clamp, is_valid_float
from .settings import PI, EPSILON, MAX_FLOAT
from .shapes import Circle, Polygon, Edge, Loop, MassData
from .fixture import Fixture
from .body import Body
from .world import World, TimeStep
from . import joints
from .joints import RevoluteJoint, DistanceJoint, WeldJoint, PrismaticJoint, RopeJoint, FrictionJoint, WheelJoint, PulleyJoint, MouseJoint
from . import controllers
from .controllers import Controller, GravityController, BuoyancyController
from .batch import BatchedWorld

__all__ = ["World", "Body", "Fixture", "Circle", "Polygon", "Edge", "Loop", "Vec2", "AABB", "Transform", "BatchedWorld", "controllers", "Controller", "GravityController", "BuoyancyController",
           "PhysicsError", "LockedError", "EmptyFixtureError", "TimeStep", "MassData",
           "scalar_cross", "clamp", "is_valid_float", "PI", "EPSILON", "MAX_FLOAT",
           "settings", "shapes", "scenes", "schema", "joints", "RevoluteJoint", "DistanceJoint", "WeldJoint", "PrismaticJoint", "RopeJoint", "FrictionJoint", "WheelJoint", "PulleyJoint", "MouseJoint"]
