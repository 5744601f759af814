"""Tuning constants of the rigid-body step.

Values follow the reference's `pypybox2d/settings.py:32-134` (Box2D 2.2 MKS tuning set).
On the device they are `constexpr` members of `b2b::K<real>` in `csrc/b2b_math.cuh`;
this module is what the host-side scene builder and the tests read.
"""
import math
import sys

PI = math.pi
EPSILON = sys.float_info.epsilon          # fp64 epsilon, as the reference (settings.py:32)
EPSILON_SQR = EPSILON ** 2
MAX_FLOAT = sys.float_info.max

MAX_MANIFOLD_POINTS = 2                   # settings.py:47
MAX_POLYGON_VERTICES = 8                  # device capacity (reference has no cap)
AABB_EXTENSION = 0.1                      # settings.py:52
AABB_MULTIPLIER = 2.0                     # settings.py:57
LINEAR_SLOP = 0.005                       # settings.py:61
ANGULAR_SLOP = 2.0 / 180.0 * PI           # settings.py:65
POLYGON_RADIUS = 2.0 * LINEAR_SLOP        # settings.py:70
VELOCITY_THRESHOLD = 1.0                  # settings.py:86
MAX_LINEAR_CORRECTION = 0.2               # settings.py:90
MAX_TRANSLATION = 2.0                     # settings.py:98
MAX_TRANSLATION_SQR = MAX_TRANSLATION ** 2
MAX_ROTATION = 0.5 * PI                   # settings.py:103
MAX_ROTATION_SQR = MAX_ROTATION ** 2
BAUMGARTE = 0.2                           # settings.py:109
TIME_TO_SLEEP = 0.5                       # settings.py:115
LINEAR_SLEEP_TOLERANCE = 0.01             # settings.py:118
LINEAR_SLEEP_TOLERANCE_SQR = LINEAR_SLEEP_TOLERANCE ** 2
ANGULAR_SLEEP_TOLERANCE = 2.0 / 180.0 * PI  # settings.py:122
ANGULAR_SLEEP_TOLERANCE_SQR = ANGULAR_SLEEP_TOLERANCE ** 2
MAX_CONDITION_NUMBER = 1000.0             # contact.py:488
RELATIVE_TOL = 0.98                       # collision.py:39
ABSOLUTE_TOL = 0.001                      # collision.py:40
