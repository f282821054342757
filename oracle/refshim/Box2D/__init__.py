"""TEST INFRASTRUCTURE: a stand-in for the pybox2d API subset the reference calls,
backed by the oracle's float32 Box2D restatement (oracle/b2lite.h through libncg_oracle.so).

It exists so the reference's own, unmodified src/car_env.py can run in the offline build container and
produce golden trajectories (oracle/gen_golden.py): everything the reference computes in Python (forces,
tyres, lap timer, reward, disable rules, termination, observation) is then the reference's arithmetic, and
only the rigid-body step underneath is this repo's restatement.  It is NOT real Box2D."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import oracle as _O  # noqa: E402

_L = _O.lib()
_f, _i, _vp = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
_L.orc_b2_create.restype = _vp
_L.orc_b2_create.argtypes = [_f] * 11
_L.orc_b2_free.argtypes = [_vp]
_L.orc_b2_add_wall.argtypes = [_vp] + [_f] * 5
_L.orc_b2_apply_force.argtypes = [_vp] + [_f] * 4
_L.orc_b2_apply_force_center.argtypes = [_vp, _f, _f]
_L.orc_b2_apply_torque.argtypes = [_vp, _f]
_L.orc_b2_set_transform.argtypes = [_vp, _f, _f, _f]
_L.orc_b2_set_linear_velocity.argtypes = [_vp, _f, _f]
_L.orc_b2_set_angular_velocity.argtypes = [_vp, _f]
_L.orc_b2_step.argtypes = [_vp, _f, _i, _i]
_L.orc_b2_get.argtypes = [_vp, ctypes.POINTER(_f)]
_L.orc_b2_raycast.argtypes = [_vp, _f, _f, _f, _f, ctypes.POINTER(_i)]
_L.orc_b2_raycast.restype = _f
_L.orc_b2_query.argtypes = [_vp, _f, _f, _f, _f, ctypes.POINTER(_i), _i]
_L.orc_b2_query.restype = _i
_L.orc_b2_test_point.argtypes = [_vp, _i, _f, _f]
_L.orc_b2_test_point.restype = _i
_L.orc_b2_wall_vertices.argtypes = [_vp, _i, ctypes.POINTER(_f)]
_L.orc_b2_events.argtypes = [_vp, ctypes.POINTER(_f), _i]
_L.orc_b2_events.restype = _i

b2_staticBody, b2_kinematicBody, b2_dynamicBody = 0, 1, 2


def _f32(x):
    return float(np.float32(x))


class b2Vec2:
    def __init__(self, x=0.0, y=0.0):
        if isinstance(x, (tuple, list, b2Vec2)):
            x, y = x[0], x[1]
        self.x, self.y = _f32(x), _f32(y)

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __len__(self):
        return 2

    def __iter__(self):
        return iter((self.x, self.y))

    @property
    def length(self):
        return float(np.sqrt(np.float32(np.float32(self.x) * np.float32(self.x) + np.float32(self.y) * np.float32(self.y))))


class b2BodyDef:
    def __init__(self):
        self.type, self.position, self.angle = b2_staticBody, (0.0, 0.0), 0.0


class _Filter:
    categoryBits, maskBits = 1, 0xFFFF


class b2FixtureDef:
    def __init__(self):
        self.shape, self.density, self.friction, self.restitution, self.filter = None, 0.0, 0.2, 0.0, _Filter()


class b2MassData:
    def __init__(self):
        self.mass, self.center, self.I = 0.0, (0.0, 0.0), 0.0


class b2AABB:
    def __init__(self):
        self.lowerBound, self.upperBound = (0.0, 0.0), (0.0, 0.0)


class b2Transform:
    def __init__(self, body, wall_index=None):
        self.body, self.wall_index = body, wall_index


class b2PolygonShape:
    def __init__(self):
        self.hx = self.hy = 0.0
        self._world, self._wall = None, None

    def SetAsBox(self, hx, hy):
        self.hx, self.hy = _f32(hx), _f32(hy)

    @property
    def vertexCount(self):
        return 4

    @property
    def vertices(self):
        return [(-self.hx, -self.hy), (self.hx, -self.hy), (self.hx, self.hy), (-self.hx, self.hy)]

    def TestPoint(self, transform, p):
        return bool(_L.orc_b2_test_point(self._world._h, self._wall, _f32(p[0]), _f32(p[1])))


def b2Mul(transform, v):
    """b2Mul(transform, local vertex) for a wall fixture: world-space vertex."""
    body = transform.body
    buf = (_f * 8)()
    _L.orc_b2_wall_vertices(body.world._h, body.wall_index, buf)
    verts = body.fixture.shape.vertices
    k = [i for i in range(4) if verts[i] == (v[0], v[1])][0]
    return b2Vec2(buf[2 * k], buf[2 * k + 1])


class b2RayCastCallback:
    def __init__(self):
        pass


class b2QueryCallback:
    def __init__(self):
        pass


class b2ContactListener:
    def __init__(self):
        pass


class _Fixture:
    def __init__(self, body, shape):
        self.body, self.shape, self.userData = body, shape, None


class _Body:
    def __init__(self, world, bdef):
        self.world, self.type = world, bdef.type
        self._pos0, self._angle0 = (bdef.position[0], bdef.position[1]), bdef.angle
        self.userData, self.fixture, self.wall_index = None, None, None
        self._pending_density = None

    # ---- fixtures / mass
    def CreateFixture(self, fdef):
        fx = _Fixture(self, fdef.shape)
        self.fixture = fx
        if self.type == b2_dynamicBody:
            self._fdef = fdef
        else:
            self.wall_index = self.world._add_wall(self, fdef)
            fdef.shape._world, fdef.shape._wall = self.world, self.wall_index
        return fx

    @property
    def massData(self):
        return None

    @massData.setter
    def massData(self, md):
        self.world._create_car(self, md)

    # ---- state (float32 values surfaced as Python floats, like SWIG does)
    def _get(self):
        buf = (_f * 8)()
        _L.orc_b2_get(self.world._h, buf)
        return [float(v) for v in buf]

    @property
    def position(self):
        if self.type != b2_dynamicBody:
            return b2Vec2(self._pos0)
        g = self._get()
        return b2Vec2(g[0], g[1])

    @position.setter
    def position(self, p):
        g = self._get()
        _L.orc_b2_set_transform(self.world._h, _f32(p[0]), _f32(p[1]), g[2])

    @property
    def angle(self):
        return self._get()[2] if self.type == b2_dynamicBody else _f32(self._angle0)

    @angle.setter
    def angle(self, a):
        g = self._get()
        _L.orc_b2_set_transform(self.world._h, g[0], g[1], _f32(a))

    @property
    def linearVelocity(self):
        g = self._get()
        return b2Vec2(g[3], g[4])

    @linearVelocity.setter
    def linearVelocity(self, v):
        _L.orc_b2_set_linear_velocity(self.world._h, _f32(v[0]), _f32(v[1]))

    @property
    def angularVelocity(self):
        return self._get()[5]

    @angularVelocity.setter
    def angularVelocity(self, w):
        _L.orc_b2_set_angular_velocity(self.world._h, _f32(w))

    @property
    def transform(self):
        return b2Transform(self, self.wall_index)

    @property
    def contacts(self):
        return []

    def GetWorldVector(self, v):
        g = self._get()
        c, s = np.float32(g[6]), np.float32(g[7])
        x, y = np.float32(v[0]), np.float32(v[1])
        return b2Vec2(c * x - s * y, s * x + c * y)

    def GetWorldPoint(self, v):
        g = self._get()
        c, s = np.float32(g[6]), np.float32(g[7])
        x, y = np.float32(v[0]), np.float32(v[1])
        return b2Vec2((c * x - s * y) + np.float32(g[0]), (s * x + c * y) + np.float32(g[1]))

    def ApplyForce(self, force, point, wake):
        _L.orc_b2_apply_force(self.world._h, _f32(force[0]), _f32(force[1]), _f32(point[0]), _f32(point[1]))

    def ApplyForceToCenter(self, force, wake):
        _L.orc_b2_apply_force_center(self.world._h, _f32(force[0]), _f32(force[1]))

    def ApplyTorque(self, torque, wake):
        _L.orc_b2_apply_torque(self.world._h, _f32(torque))


class _WorldManifold:
    def __init__(self, n):
        self.points, self.normal = ((0.0, 0.0), (0.0, 0.0)), (float(n[0]), float(n[1]))


class _Contact:
    def __init__(self, car, wall, normal=(0.0, 0.0)):
        self.fixtureA, self.fixtureB = car.fixture, wall.fixture
        self.worldManifold = _WorldManifold(normal)


class _Impulse:
    def __init__(self, vals):
        self.normalImpulses = vals


class b2World:
    def __init__(self, gravity=(0, 0), doSleep=True):
        self._h, self.car, self.walls, self.contactListener = None, None, [], None
        self._pending_walls = []

    def __del__(self):
        try:
            if self._h:
                _L.orc_b2_free(self._h)
        except Exception:
            pass

    def CreateBody(self, bdef):
        return _Body(self, bdef)

    def _create_car(self, body, md):
        fd = body._fdef
        # wall material is fixed by the reference's constants (physics.py:6-9)
        from src.constants import BOX2D_WALL_FRICTION, BOX2D_WALL_RESTITUTION
        self._h = _L.orc_b2_create(_f32(body._pos0[0]), _f32(body._pos0[1]), _f32(body._angle0), fd.shape.hx, fd.shape.hy, _f32(md.mass),
                                   _f32(md.I), _f32(fd.friction), _f32(fd.restitution), _f32(BOX2D_WALL_FRICTION), _f32(BOX2D_WALL_RESTITUTION))
        self.car = body

    def _add_wall(self, body, fdef):
        _L.orc_b2_add_wall(self._h, _f32(body._pos0[0]), _f32(body._pos0[1]), _f32(body._angle0), fdef.shape.hx, fdef.shape.hy)
        self.walls.append(body)
        return len(self.walls) - 1

    @property
    def bodies(self):
        return [self.car] + self.walls

    def IsLocked(self):
        return False

    def Step(self, dt, vel_iters, pos_iters):
        _L.orc_b2_step(self._h, _f32(dt), vel_iters, pos_iters)
        buf = (_f * (7 * 256))()
        n = _L.orc_b2_events(self._h, buf, 256)
        lis = self.contactListener
        if lis is None:
            return
        for k in range(n):
            e = buf[7 * k: 7 * k + 7]
            wall = self.walls[int(e[1])]
            if int(e[0]) == 0:
                lis.BeginContact(_Contact(self.car, wall, (e[2], e[3])))
            elif int(e[0]) == 1:
                lis.EndContact(_Contact(self.car, wall))
            else:
                lis.PostSolve(_Contact(self.car, wall), _Impulse([float(e[5 + j]) for j in range(int(e[4]))]))

    def RayCast(self, callback, p1, p2):
        wall = _i(-1)
        fr = _L.orc_b2_raycast(self._h, _f32(p1[0]), _f32(p1[1]), _f32(p2[0]), _f32(p2[1]), ctypes.byref(wall))
        if wall.value >= 0:
            x = (1.0 - fr) * p1[0] + fr * p2[0]
            y = (1.0 - fr) * p1[1] + fr * p2[1]
            callback.ReportFixture(self.walls[wall.value].fixture, (x, y), (0.0, 0.0), float(fr))

    def QueryAABB(self, callback, aabb):
        out = (_i * 64)()
        n = _L.orc_b2_query(self._h, _f32(aabb.lowerBound[0]), _f32(aabb.lowerBound[1]), _f32(aabb.upperBound[0]), _f32(aabb.upperBound[1]),
                            out, 64)
        for k in range(n):
            if not callback.ReportFixture(self.walls[out[k]].fixture):
                break

b2Body = _Body
b2Fixture = _Fixture
b2Contact = _Contact
