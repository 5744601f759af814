"""Controllers: the drop-in surface of `pypybox2d/controllers.py`.

A controller is host code in the reference too: once per step, between the narrowphase and the island solver (`world.py:477-479`),
it applies forces to its bodies from their current state.  Here `World.step` runs the narrowphase on the device first (a step with
dt = 0, `world.py:338-349`: contacts updated, bodies woken by touching changes), mirrors the awake flags back, lets the controllers
apply their forces to the `Body` objects with the reference's arithmetic, uploads them and runs the step.  Worlds without controllers
never take this path; large batches apply forces through `b2b_step_host`.
"""
import math

from .common import Vec2
from .settings import EPSILON, PI
from . import shapes as _shapes


class Controller(object):
    """controllers.py:25-78"""

    def __init__(self, world, bodies=[]):
        self._bodies = []
        self._world = world
        for body in bodies:
            self.add_body(body)

    bodies = property(lambda self: list(self._bodies))
    world = property(lambda self: self._world)

    def step(self, timestep):
        raise NotImplementedError

    def add_body(self, body):
        if body in self._bodies:
            raise ValueError('Body already in controller')
        elif not body:
            raise ValueError('Body unset')
        elif body._world != self._world:
            raise ValueError('Body not in same world')
        self._bodies.append(body)
        body._controllers.append(self)

    def remove_body(self, body):
        if not body:
            raise ValueError('Body unset')
        elif body not in self._bodies:
            raise ValueError('Body not in controller')
        self._bodies.remove(body)
        body._controllers.remove(self)

    def clear(self):
        for body in list(self._bodies):
            self.remove_body(body)


def _xf_mul(xf, x, y):
    px, py, c, s = xf
    return px + c * x + (-s) * y, py + s * x + c * y


def _submerged_circle(shape, normal, offset, xf):
    """Circle.compute_submerged_area (shapes.py:219-237)"""
    px, py = _xf_mul(xf, shape.position.x, shape.position.y)
    radius = shape.radius
    l = -((normal.x * px + normal.y * py) - offset)
    if l < -radius + EPSILON:
        return 0.0, (px, py)
    elif l > radius:
        return PI * radius ** 2, (px, py)
    r2 = radius ** 2
    l2 = l ** 2
    area = r2 * (math.asin(l / radius) + PI / 2.0) + l * math.sqrt(r2 - l2)
    com = -2.0 / 3.0 * pow(r2 - l2, 1.5) / area
    return area, (px + normal.x * com, py + normal.y * com)


def _submerged_polygon(shape, normal, offset, xf, density):
    """Polygon.compute_submerged_area (shapes.py:538-616)"""
    vertices = shape._vertices
    n = len(vertices)
    px, py, c, s = xf
    nlx, nly = c * normal.x + s * normal.y, (-s) * normal.x + c * normal.y          # xf._rotation.mul_t(normal)
    offset_l = offset - (normal.x * px + normal.y * py)
    depths = [(nlx * v.x + nly * v.y) - offset_l for v in vertices]
    last_submerged = False
    into_index = outof_index = None
    dive_count = 0
    for i, depth in enumerate(depths):
        is_submerged = depth < EPSILON
        if i > 0:
            if is_submerged:
                if not last_submerged:
                    into_index = i - 1
                    dive_count += 1
            else:
                if last_submerged:
                    outof_index = i - 1
                    dive_count += 1
        last_submerged = is_submerged
    if dive_count == 0:
        md = shape.compute_mass(density)
        center = _xf_mul(xf, md.center.x, md.center.y)
        return (md.mass / density if last_submerged else 0.0), center
    elif dive_count == 1:
        if into_index is None:
            into_index = n - 1
        else:
            outof_index = n - 1
    into_index2 = (into_index + 1) % n
    outof_index2 = (outof_index + 1) % n
    into_lambda = -depths[into_index] / (depths[into_index2] - depths[into_index])
    outof_lambda = -depths[outof_index] / (depths[outof_index2] - depths[outof_index])

    def lerp(a, b, t):
        # vertices[a] * (1.0 - t) + vertices[b] * t, Vec2 * number = Vec2(number * x, number * y)
        k = 1.0 - t
        return (k * vertices[a].x + t * vertices[b].x, k * vertices[a].y + t * vertices[b].y)

    into_vec = lerp(into_index, into_index2, into_lambda)
    outof_vec = lerp(outof_index, outof_index2, outof_lambda)
    area = 0
    cx, cy = 0.0, 0.0
    p2 = (vertices[into_index2].x, vertices[into_index2].y)
    i = into_index2
    while i != outof_index2:
        i = (i + 1) % n
        p3 = outof_vec if i == outof_index2 else (vertices[i].x, vertices[i].y)
        e1 = (p2[0] - into_vec[0], p2[1] - into_vec[1])
        e2 = (p3[0] - into_vec[0], p3[1] - into_vec[1])
        d = e1[0] * e2[1] - e1[1] * e2[0]
        triangle_area = 0.5 * d
        area += triangle_area
        # center += triangle_area * (into_vec + p2 + p3) / 3.0
        sx, sy = (into_vec[0] + p2[0]) + p3[0], (into_vec[1] + p2[1]) + p3[1]
        cx += (triangle_area * sx) / 3.0
        cy += (triangle_area * sy) / 3.0
        p2 = p3
    k = 1.0 / area
    cx, cy = k * cx, k * cy
    return area, _xf_mul(xf, cx, cy)


def _submerged(shape, normal, offset, xf, density):
    if shape._type == _shapes.CIRCLE:
        return _submerged_circle(shape, normal, offset, xf)
    if shape._type == _shapes.POLYGON:
        return _submerged_polygon(shape, normal, offset, xf, density)
    if shape._type == _shapes.EDGE:
        return 0.0, (0.0, 0.0)                      # shapes.py:751-753
    raise NotImplementedError("compute_submerged_area of a Loop (marked 'likely wrong' in the reference, shapes.py:866-880)")


class BuoyancyController(Controller):
    """controllers.py:81-228: buoyancy, linear and angular drag of a fluid filling the half plane normal . p < offset."""

    def __init__(self, world, bodies=[], normal=(0, 1), offset=0, density=0, velocity=(0, 0), linear_drag=0, angular_drag=0,
                 use_density=False, gravity=None):
        Controller.__init__(self, world, bodies)
        self.normal = Vec2(*normal)
        self.offset = float(offset)
        self.density = float(density)
        self.velocity = Vec2(*velocity)
        self.linear_drag = float(linear_drag)
        self.angular_drag = float(angular_drag)
        if use_density:
            raise NotImplementedError("use_density=True reads Shape.density, which the reference's shapes do not have (controllers.py:200)")
        self.use_density = False
        self.gravity = None if gravity is None else Vec2(*gravity)

    def step(self, timestep):
        gravity = self.gravity if self.gravity else self._world._gravity
        normal, offset = self.normal, self.offset
        for body in [b for b in self._bodies if b.awake]:
            acx = acy = mcx = mcy = 0.0
            area = mass = 0.0
            xf = body._xf
            for fixture in body.fixtures:
                shape_density = 1.0
                s_area, sc = _submerged(fixture.shape, normal, offset, xf, shape_density)
                area += s_area
                acx += s_area * sc[0]
                acy += s_area * sc[1]
                mass += s_area * shape_density
                mcx += (s_area * shape_density) * sc[0]
                mcy += (s_area * shape_density) * sc[1]
            if area >= EPSILON:
                acx, acy = acx / area, acy / area
                mcx, mcy = mcx / mass, mcy / mass
                k = -self.density * area
                body.apply_force((k * gravity.x, k * gravity.y), (mcx, mcy))
                lv = body.get_linear_velocity_from_world_point((acx, acy))
                kd = -self.linear_drag * area
                body.apply_force((kd * (lv.x - self.velocity.x), kd * (lv.y - self.velocity.y)), (acx, acy))
                body.apply_torque(-body.inertia / body.mass * area * body._angular_velocity * self.angular_drag)


class GravityController(Controller):
    """controllers.py:231-280: simplified gravity between every (ordered) pair of its bodies."""

    def __init__(self, world, bodies=[], G=1.0, inv_sqr=True):
        Controller.__init__(self, world, bodies)
        self.G = float(G)
        self.inv_sqr = bool(inv_sqr)

    def step(self, timestep):
        G, inv_sqr = self.G, self.inv_sqr
        for body1 in self._bodies:
            for body2 in self._bodies:
                if body1 is body2:
                    continue
                c1, c2 = body1._c, body2._c
                dx, dy = c2.x - c1.x, c2.y - c1.y
                r2 = dx ** 2 + dy ** 2
                k = G / r2 * body1.mass * body2.mass
                fx, fy = k * dx, k * dy
                if inv_sqr:
                    r = math.sqrt(r2)
                    fx, fy = fx / r, fy / r
                body1.apply_force((fx, fy), (body1._c.x, body1._c.y))
                body2.apply_force((-fx, -fy), (body2._c.x, body2._c.y))
