"""Collision shapes of the drop-in API (host side, fp64).

Scene construction stays on the host in Python floats so that masses, centroids and
edge normals are bit-identical to what the reference computes
(`pypybox2d/shapes.py:128-236` Circle, `:238-616` Polygon, `:618-778` Edge).  Only the
pieces the step path needs are provided: geometry, `compute_mass`, `compute_aabb`.
Ray casts, point tests, buoyancy areas and `Loop` are out of scope.
"""
import math

from .common import Vec2
from .settings import EPSILON, EPSILON_SQR, PI, POLYGON_RADIUS, MAX_POLYGON_VERTICES

# integer type ids, same numbering as shapes.py:884-889
CIRCLE, POLYGON, LOOP, EDGE = 0, 1, 2, 3


class MassData(object):
    def __init__(self, mass=0.0, center=(0.0, 0.0), I=0.0):
        self.mass = mass
        self.center = Vec2(*center)
        self.I = I

    def __repr__(self):
        return "MassData(mass=%g, center=%s, I=%g)" % (self.mass, self.center, self.I)


class Shape(object):
    radius = 0.0
    _type = -1
    child_count = 1

    def compute_aabb(self, xf):
        raise NotImplementedError

    def compute_mass(self, density):
        raise NotImplementedError


def _xf_mul(xf, v):
    px, py, c, s = xf
    return (px + c * v[0] + (-s) * v[1], py + s * v[0] + c * v[1])


class Circle(Shape):
    _type = CIRCLE

    def __init__(self, radius=0.2, position=(0, 0)):
        self.position = Vec2(*position)
        self.radius = radius

    def __copy__(self):
        return Circle(self.radius, self.position)

    def __repr__(self):
        return "Circle(radius=%g, position=%s)" % (self.radius, self.position)

    def compute_aabb(self, xf):
        # shapes.py:178-189: p = xf.position + R*position, +- radius
        px, py, c, s = xf
        x = px + (c * self.position.x + (-s) * self.position.y)
        y = py + (s * self.position.x + c * self.position.y)
        r = self.radius
        return (x - r, y - r, x + r, y + r)

    def compute_mass(self, density):
        # shapes.py:191-201
        mass = density * PI * self.radius ** 2
        I = mass * (0.5 * self.radius ** 2 + self.position.dot(self.position))
        return MassData(mass, self.position, I)

    def _key(self):
        return (CIRCLE, float(self.radius), self.position.x, self.position.y)


def _edge_normal(ex, ey):
    # edge.cross(1.0).normalized  (shapes.py:450-455, common.py:204-222)
    nx, ny = 1.0 * ey, -1.0 * ex
    length = math.sqrt(nx ** 2 + ny ** 2)
    if length < EPSILON:
        return (nx, ny)
    inv = 1.0 / length
    return (nx * inv, ny * inv)


class Polygon(Shape):
    """Convex polygon, CCW, 3..8 vertices.  `box=(hx, hy[, center[, angle]])` builds a box."""
    _type = POLYGON
    CHECK_VERTICES = True

    def __init__(self, vertices=None, box=None):
        self.radius = POLYGON_RADIUS
        self._vertices = []
        self._normals = []
        # The reference never updates `_centroid`; the colliders read (0,0).
        self._centroid = Vec2()
        if box:
            self.set_as_box(*box)
        elif vertices is not None:
            self.vertices = [Vec2(*v) for v in vertices]

    def __copy__(self):
        return Polygon(self._vertices)

    def __repr__(self):
        return "Polygon(%s)" % self._vertices

    @property
    def vertices(self):
        return self._vertices

    @vertices.setter
    def vertices(self, vertices):
        verts = [(float(v[0]), float(v[1])) for v in vertices]
        n = len(verts)
        if n < 3:
            raise ValueError("Polygons are >= 3 vertices")
        if n > MAX_POLYGON_VERTICES:
            raise ValueError("Polygons are <= %d vertices on the batched path" % MAX_POLYGON_VERTICES)
        normals = []
        for i in range(n):
            ex = verts[(i + 1) % n][0] - verts[i][0]
            ey = verts[(i + 1) % n][1] - verts[i][1]
            if ex ** 2 + ey ** 2 <= EPSILON_SQR:
                raise ValueError("Edges are too short")
            normals.append(_edge_normal(ex, ey))
        if Polygon.CHECK_VERTICES:
            # convex, interior on the left of every edge (shapes.py:457-473)
            for i1 in range(n):
                i2 = (i1 + 1) % n
                ex = verts[i2][0] - verts[i1][0]
                ey = verts[i2][1] - verts[i1][1]
                for j in range(n):
                    if j == i1 or j == i2:
                        continue
                    rx = verts[j][0] - verts[i1][0]
                    ry = verts[j][1] - verts[i1][1]
                    if ex * ry - ey * rx <= 0.0:
                        raise ValueError("Polygon must be convex with CCW winding")
        self._normals = [Vec2(*nrm) for nrm in normals]
        self._vertices = [Vec2(*v) for v in verts]

    def set_as_box(self, hx, hy, center=(0, 0), angle=0.0):
        # shapes.py:484-501: corner list then rigid transform by (center, angle);
        # normals are re-derived from the edges, which is what the reference's fixture copy does.
        corners = [(-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy)]
        if center is not None:
            xf = (float(center[0]), float(center[1]), math.cos(angle), math.sin(angle))
            corners = [_xf_mul(xf, v) for v in corners]
        self.vertices = corners

    def compute_aabb(self, xf):
        # shapes.py:316-334
        lx, ly = _xf_mul(xf, self._vertices[0])
        ux, uy = lx, ly
        for v in self._vertices:
            x, y = _xf_mul(xf, v)
            lx = min(lx, x)
            ly = min(ly, y)
            ux = max(ux, x)
            uy = max(uy, y)
        r = self.radius
        return (lx - r, ly - r, ux + r, uy + r)

    def compute_mass(self, density):
        # triangle fan about the vertex mean, shapes.py:366-419 (operation order kept)
        verts = self._vertices
        n = len(verts)
        sx = sy = None
        for v in verts:
            if sx is None:
                sx, sy = v.x, v.y
            else:
                sx, sy = sx + v.x, sy + v.y
        sx, sy = sx / n, sy / n
        inv3 = 1.0 / 3.0
        cx = cy = 0.0
        area = 0.0
        I = 0.0
        for i in range(n):
            v = verts[i]
            w = verts[(i + 1) % n]
            e1x, e1y = v.x - sx, v.y - sy
            e2x, e2y = w.x - sx, w.y - sy
            D = e1x * e2y - e1y * e2x
            tri = 0.5 * D
            area += tri
            k = tri * inv3
            cx, cy = cx + k * (e1x + e2x), cy + k * (e1y + e2y)
            intx2 = e1x * e1x + e2x * e1x + e2x * e2x
            inty2 = e1y * e1y + e2y * e1y + e2y * e2y
            I += (0.25 * inv3 * D) * (intx2 + inty2)
        mass = density * area
        assert area > EPSILON
        cx, cy = cx / area, cy / area
        mcx, mcy = cx + sx, cy + sy
        Iout = density * I
        Iout += mass * ((mcx ** 2 + mcy ** 2) - (cx ** 2 + cy ** 2))
        return MassData(mass, (mcx, mcy), Iout)

    def _key(self):
        return (POLYGON, float(self.radius)) + tuple(c for v in self._vertices for c in (v.x, v.y))


class Edge(Shape):
    """Line segment v1->v2 with optional ghost vertices v0, v3 (shapes.py:618-778)."""
    _type = EDGE

    def __init__(self, v1=None, v2=None, v0=None, v3=None):
        self.radius = POLYGON_RADIUS
        self._vertex1 = Vec2(*v1)
        self._vertex2 = Vec2(*v2)
        # the reference tests truthiness; any non-None 2-sequence counts (shapes.py:634-642)
        self._vertex0 = Vec2(*v0) if v0 is not None else None
        self._vertex3 = Vec2(*v3) if v3 is not None else None

    def __copy__(self):
        return Edge(self._vertex1, self._vertex2, self._vertex0, self._vertex3)

    def __repr__(self):
        return "Edge(vertex1=%s, vertex2=%s, [vertex0=%s, vertex3=%s])" % (
            self._vertex1, self._vertex2, self._vertex0, self._vertex3)

    vertex1 = property(lambda self: Vec2(*self._vertex1))
    vertex2 = property(lambda self: Vec2(*self._vertex2))
    vertex0 = property(lambda self: None if self._vertex0 is None else Vec2(*self._vertex0))
    vertex3 = property(lambda self: None if self._vertex3 is None else Vec2(*self._vertex3))

    _loop_child = False          # a child edge of a Loop: its AABB has no radius margin (shapes.py:822-837)

    def compute_aabb(self, xf):
        # shapes.py:702-718
        x1, y1 = _xf_mul(xf, self._vertex1)
        x2, y2 = _xf_mul(xf, self._vertex2)
        r = 0.0 if self._loop_child else self.radius
        return (min(x1, x2) - r, min(y1, y2) - r, max(x1, x2) + r, max(y1, y2) + r)

    def compute_mass(self, density):
        c = 0.5 * (self._vertex1 + self._vertex2)
        return MassData(0.0, c, 0.0)

    def _key(self):
        v0 = self._vertex0
        v3 = self._vertex3
        return (EDGE, float(self.radius), self._vertex1.x, self._vertex1.y, self._vertex2.x, self._vertex2.y,
                None if v0 is None else (v0.x, v0.y), None if v3 is None else (v3.x, v3.y), bool(self._loop_child))


class Loop(Shape):
    """A closed sequence of line segments (shapes.py:780-880): one broadphase proxy and one contact slot per child edge, each
    child colliding as an Edge with its two neighbours as ghost vertices (Loop.child_edge); zero mass; a child's AABB is the
    segment's box WITHOUT the radius margin (shapes.py:822-837, unlike Edge.compute_aabb)."""
    _type = LOOP

    def __init__(self, vertices, *args):
        if len(vertices) == 2 and isinstance(vertices[0], (int, float)):
            vertices = [vertices]
            vertices.extend(args)
        if len(vertices) < 2:
            raise ValueError('Must be >= 2 vertices')
        self.radius = POLYGON_RADIUS
        self._vertices = [Vec2(*v) for v in vertices]

    def __copy__(self):
        return Loop(self._vertices)

    def __repr__(self):
        return 'Loop(vertices=%s)' % self._vertices

    @property
    def child_count(self):
        return len(self._vertices)

    @property
    def vertices(self):
        return [Vec2(*v) for v in self._vertices]

    def get_vertex(self, index):
        return Vec2(*self._vertices[index])

    def child_edge(self, index):
        count = len(self._vertices)
        assert 0 <= index < count
        e = Edge(self._vertices[index], self._vertices[(index + 1) % count], self._vertices[(index - 1) % count],
                 self._vertices[(index + 2) % count])
        e._loop_child = True
        return e

    def test_point(self, xf, p):
        return False

    def compute_aabb(self, xf, child_index=0):
        x1, y1 = _xf_mul(xf, self._vertices[child_index])
        x2, y2 = _xf_mul(xf, self._vertices[(child_index + 1) % len(self._vertices)])
        return (min(x1, x2), min(y1, y2), max(x1, x2), max(y1, y2))

    def compute_mass(self, density):
        return MassData()

    def _key(self):
        return (LOOP, float(self.radius)) + tuple(c for v in self._vertices for c in (v.x, v.y))
