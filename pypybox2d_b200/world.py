"""`World`: the drop-in surface of the reference (`pypybox2d/world.py:64-91,183-233,307-369`).

Scene construction is host-side Python; `step(dt, vel_iters, pos_iters)` packs the world into
device structure-of-arrays buffers (once, or again after a host-side mutation), runs the CUDA
step through the C-ABI (`include/b2b.h`) and mirrors the body state back into the `Body`
objects.  A single `World` is a batch of one; large batches use `BatchedWorld` directly
(`BatchedWorld.replicate(world, n, ...)`), which never touches per-body Python on the step path.

Revolute, distance, weld, prismatic, rope, friction, wheel, pulley and mouse joints (`joints.py`) are solved inside the island solver kernels.  Unsupported on this path
(raise `NotImplementedError`): the gear joint, the `pre_solve` listener,
custom contact filters and sub-stepping.  Continuous collision
(`continuous_physics=True`, the reference's default: `World.solve_toi`, world.py:594-821) runs on the device.
"""
from copy import copy

from .body import Body
from . import schema
from . import joints as _joints
from .common import Vec2, LockedError
from .proxytree import ProxyTree
from .settings import AABB_EXTENSION


class TimeStep(object):
    """world.py:53-62"""
    __slots__ = ["dt", "inv_dt", "dt_ratio", "vel_iters", "pos_iters", "warm_starting"]

    def __init__(self, dt=0.0, inv_dt=0.0, dt_ratio=0.0, vel_iters=0, pos_iters=0, warm_starting=True):
        self.dt = float(dt)
        self.inv_dt = float(inv_dt)
        self.dt_ratio = float(dt_ratio)
        self.vel_iters = int(vel_iters)
        self.pos_iters = int(pos_iters)
        self.warm_starting = bool(warm_starting)


class ContactView(object):
    """Read-only record of one device contact row (subset of `pypybox2d/contact.py:49-201`)."""
    __slots__ = ["fixture_a", "fixture_b", "touching", "enabled", "friction", "restitution",
                 "manifold_type", "point_count", "local_normal", "local_point", "points"]

    @property
    def body_a(self):
        return self.fixture_a.body

    @property
    def body_b(self):
        return self.fixture_b.body


class _BroadPhaseView(object):
    """`world.contact_manager.broadphase` as far as a caller can observe it (broadphase.py:78-81, 137-150): the proxy count and
    the tree observables, from the host replay of the reference's tree topology (`proxytree.ProxyTree`)."""
    def __init__(self, world):
        self._world = world

    @property
    def proxy_count(self):
        """The reference's counter, not the number of leaves: `destroy_body` walks the body's fixtures twice
        (body.py:330-340) and `BroadPhase.destroy_proxy` decrements before it looks at the proxy (broadphase.py:53-59), so
        every proxy of a destroyed body is subtracted twice."""
        return self._world._proxy_count

    @property
    def tree_height(self):
        return self._world._tree.tree_height

    @property
    def tree_balance(self):
        return self._world._tree.tree_balance

    @property
    def tree_quality(self):
        return self._world._tree.tree_quality


class _ContactManagerView(object):
    def __init__(self, world):
        self._world = world
        self.broadphase = _BroadPhaseView(world)
        self._post_solve = None

    @property
    def contacts(self):
        return self._world._contact_views()

    # Contact listeners (contact_manager.py:46-100).  begin_contact / end_contact are delivered after the step that caused
    # them, in the reference's order (the narrowphase walks the contact list: contact.py:263-266, contact_manager.py:167-168),
    # from the difference between the device contact lists before and after the step; post_solve is delivered after the step from the island
    # lists and the stored impulses (World._deliver_post_solve); pre_solve would have to run between the narrowphase and the
    # solver kernels of every step and is not supported.
    def __init_listeners(self):
        self._begin_contact = None
        self._end_contact = None

    @property
    def begin_contact(self):
        return getattr(self, "_begin_contact", None)

    @begin_contact.setter
    def begin_contact(self, fcn):
        self._begin_contact = fcn
        self._world._listeners_changed()

    @property
    def end_contact(self):
        return getattr(self, "_end_contact", None)

    @end_contact.setter
    def end_contact(self, fcn):
        self._end_contact = fcn
        self._world._listeners_changed()

    @property
    def pre_solve(self):
        return None

    @pre_solve.setter
    def pre_solve(self, fcn):
        if fcn is not None:
            raise NotImplementedError("pre_solve callbacks cannot run inside the device step (SURVEY §8b)")

    @property
    def post_solve(self):
        return self._post_solve

    @post_solve.setter
    def post_solve(self, fcn):
        """`post_solve(contact, impulses)` (contact_manager.py:82-96): what `Island.report` (island.py:245-252) hands out after an
        island's solve -- the island's contacts in island order with the accumulated (normal, tangent) impulse of every manifold
        point.  Nothing in it feeds back into the step, so it is delivered after the device step from the island lists and the
        stored impulses (`World._deliver_post_solve`)."""
        self._post_solve = fcn

    @property
    def contact_listener(self):
        return [self.begin_contact, self.end_contact, None, self.post_solve]


class World(object):
    def __init__(self, gravity=(0, -10), do_sleep=True, warm_starting=True,
                 continuous_physics=True, sub_stepping=False, auto_clear_forces=True,
                 joint_destruction_listener=None, fixture_destruction_listener=None,
                 contact_filter=None, real="f32", device=None):
        if contact_filter is not None:
            raise NotImplementedError("custom contact filters cannot run inside the device step")
        if sub_stepping:
            raise NotImplementedError("sub_stepping (single-stepping through TOI events, a testbed debugging aid) is not supported")
        self._gravity = Vec2(*gravity)
        self.bodies = []
        self.joints = []
        self.controllers = []
        self._warm_starting = bool(warm_starting)
        self._continuous_physics = bool(continuous_physics)
        self._sub_stepping = False
        self._allow_sleep = bool(do_sleep)
        self._auto_clear_forces = bool(auto_clear_forces)
        self.joint_destruction_listener = joint_destruction_listener
        self.fixture_destruction_listener = fixture_destruction_listener
        self.contact_manager = _ContactManagerView(self)
        self._inv_dt0 = 0.0
        self._locked = False
        self._stepped = False
        self._host_dirty = True
        self._structure_dirty = False
        self._packed_fixtures = []
        self._packed_keys = []
        self._prev_contacts = None
        self._batch = None
        self._real = real
        self._device = device
        # the reference's broadphase tree topology (row a19), replayed for tree_height / tree_balance / tree_quality
        self._tree = ProxyTree()
        self._tree_leaf = {}            # (id(fixture), child) -> leaf slot
        self._proxy_count = 0
        self._tree_fat = None           # the device's fat boxes as of the last replay, rows in `_packed_keys` order
        self._moved_proxies = {}        # proxies of bodies whose transform the host set: key -> (tight box, re-inserted?)
        self._moved_body = None
        self._activated = set()         # proxies of bodies activated since the world was packed: fresh boxes, in the move buffer
        self._flagged_contacts = set()  # contacts flagged for re-filtering (contact.py:77): (proxy key a, proxy key b)
        self._touched_proxies = set()   # proxies put back into the move buffer by Fixture._refilter (fixture.py:283-285)

    # -- properties (world.py:93-181) ------------------------------------------------------
    @property
    def gravity(self):
        return Vec2(*self._gravity)

    @gravity.setter
    def gravity(self, g):
        self._gravity = Vec2(*g)
        self._host_dirty = True

    @property
    def locked(self):
        return self._locked

    warm_starting = property(lambda self: self._warm_starting)
    allow_sleep = property(lambda self: self._allow_sleep)
    auto_clear_forces = property(lambda self: self._auto_clear_forces)

    @property
    def continuous_physics(self):
        return self._continuous_physics

    @continuous_physics.setter
    def continuous_physics(self, v):
        self._continuous_physics = bool(v)
        if self._batch is not None:
            self._batch.set_continuous(self._continuous_physics)

    @property
    def contact_filter(self):
        return self.default_contact_filter

    # -- body factories (world.py:183-233) -------------------------------------------------
    def add_body(self, body):
        if self._locked:
            raise LockedError("Cannot create a body while simulating")
        self.bodies.append(body)
        body._world = self
        for fixture in body.fixtures:          # Body._attached_to_world, body.py:292-297
            self._tree_create(body, fixture)
        self._host_dirty = True
        if self._stepped:
            self._structure_dirty = True       # the device arrays are re-packed before the next step (_rebuild_batch)
        return body

    def create_body(self, body=None, **kwargs):
        if self._locked:
            raise LockedError("Cannot create a body while simulating")
        if body is not None:
            body = copy(body)                  # world.py:219-220; kwargs (the type of create_*_body included) are ignored, as there
        else:
            body = Body(**kwargs)
        return self.add_body(body)

    def create_dynamic_body(self, body=None, **kwargs):
        return self.create_body(body, type=Body.DYNAMIC, **kwargs)

    def create_static_body(self, body=None, **kwargs):
        return self.create_body(body, type=Body.STATIC, **kwargs)

    def create_kinematic_body(self, body=None, **kwargs):
        return self.create_body(body, type=Body.KINEMATIC, **kwargs)

    def destroy_body(self, body):
        """world.py:235-251: the body leaves the world with its fixtures' proxies and its contacts (Body._detached_from_world,
        body.py:310-346); it stays a valid object and can be added to a world again."""
        if self._locked:
            raise LockedError("Cannot remove a body while simulating")
        if body not in self.bodies:
            raise ValueError("Body not in this world")
        for joint in list(body.joints):          # Body._detached_from_world, body.py:313-320
            if self.joint_destruction_listener:
                self.joint_destruction_listener(joint)
            self.destroy_joint(joint)
        for fixture in body.fixtures:            # body.py:330-335
            if self.fixture_destruction_listener:
                self.fixture_destruction_listener(fixture)
            n = self._tree_destroy(fixture)
            self._proxy_count -= n               # ... and 337-340: the second walk only moves the counter
        self.bodies.remove(body)
        body._world = None
        self._host_dirty = True
        self._structure_dirty = True

    # -- joints (world.py:252-305, 823-915) ------------------------------------------------------------------------------
    def add_joint(self, joint, duplicate=False):
        """world.py:252-283: the joint joins World.joints and its bodies' lists (body_a's first); creating a joint does not
        wake the bodies.  On a world that is already stepping the device arrays are re-packed before the next step; contacts
        between two bodies that a joint with collide_connected=False now connects are destroyed by that step's narrowphase
        (the reference flags them for filtering, contact_manager.py:199-207) -- here they are dropped by the re-pack."""
        if self._locked:
            raise LockedError('Cannot create a joint while simulating')
        if duplicate:
            joint = copy(joint)
        self.joints.append(joint)
        joint._world = self
        if joint._body_a is not None:
            joint._body_a.joints.append(joint)
        if joint._body_b is not None:
            joint._body_b.joints.append(joint)
        self._host_dirty = True
        if self._stepped:
            self._structure_dirty = True
        return joint

    def destroy_joint(self, joint):
        """world.py:285-305"""
        if self._locked:
            raise LockedError('Cannot destroy a joint while simulating')
        if joint not in self.joints:
            raise ValueError('Joint not in world')
        self.joints.remove(joint)
        joint._world = None
        for body in joint.bodies:
            if body is not None:
                body.awake = True
                body.joints.remove(joint)
        self._host_dirty = True
        self._structure_dirty = True

    def create_revolute_joint(self, *args, **kwargs):
        return self.add_joint(_joints.RevoluteJoint(*args, **kwargs))

    def create_distance_joint(self, *args, **kwargs):
        return self.add_joint(_joints.DistanceJoint(*args, **kwargs))

    def create_weld_joint(self, *args, **kwargs):
        return self.add_joint(_joints.WeldJoint(*args, **kwargs))

    def create_prismatic_joint(self, *args, **kwargs):
        return self.add_joint(_joints.PrismaticJoint(*args, **kwargs))

    def create_rope_joint(self, *args, **kwargs):
        return self.add_joint(_joints.RopeJoint(*args, **kwargs))

    def create_friction_joint(self, *args, **kwargs):
        return self.add_joint(_joints.FrictionJoint(*args, **kwargs))

    def create_wheel_joint(self, *args, **kwargs):
        return self.add_joint(_joints.WheelJoint(*args, **kwargs))

    def create_pulley_joint(self, *args, **kwargs):
        return self.add_joint(_joints.PulleyJoint(*args, **kwargs))

    def create_mouse_joint(self, *args, **kwargs):
        return self.add_joint(_joints.MouseJoint(*args, **kwargs))

    def create_gear_joint(self, *args, **kwargs):
        raise NotImplementedError("gear joints do not run on the batched step path (SURVEY §8f #3: revolute, distance, weld, prismatic, "
                                  "rope, friction, wheel, pulley and mouse joints do)")

    # -- controllers (world.py:923-956; controllers.py) -----------------------------------------------------------------------
    def add_controller(self, controller):
        self.controllers.append(controller)
        return controller

    def destroy_controller(self, controller):
        controller.clear()
        self.controllers.remove(controller)

    def create_gravity_controller(self, *args, **kwargs):
        from . import controllers
        return self.add_controller(controllers.GravityController(self, *args, **kwargs))

    def create_buoyancy_controller(self, *args, **kwargs):
        from . import controllers
        return self.add_controller(controllers.BuoyancyController(self, *args, **kwargs))

    def _fixture_attached(self, body, fixture):
        self._tree_create(body, fixture)         # body.py:203-205
        self._host_dirty = True
        if self._stepped:
            self._structure_dirty = True

    def _fixture_detached(self, body, fixture):
        self._tree_destroy(fixture)              # body.py:281-283
        self._host_dirty = True
        self._structure_dirty = True

    # -- the broadphase tree's topology (row a19; proxytree.py) ---------------------------------------------------------------
    def _tree_create(self, body, fixture):
        """fixture.py:250-261: one leaf per child of the fixture's shape, fat box = the child's box at the body's transform
        grown by AABB_EXTENSION (dynamictree.py:121-126); an inactive body's fixtures have no proxies."""
        if not body._active:
            return
        shape = fixture._shape
        for child in range(shape.child_count):
            lx, ly, ux, uy = shape.compute_aabb(body._xf, child) if shape.child_count > 1 else shape.compute_aabb(body._xf)
            self._tree_leaf[(id(fixture), child)] = self._tree.create_leaf(
                (lx - AABB_EXTENSION, ly - AABB_EXTENSION, ux + AABB_EXTENSION, uy + AABB_EXTENSION))
            self._proxy_count += 1
        self._tree_fat = None

    def _tree_destroy(self, fixture):
        n = 0
        for child in range(fixture._shape.child_count):
            slot = self._tree_leaf.pop((id(fixture), child), None)
            if slot is not None:
                self._tree.destroy_leaf(slot)
                self._proxy_count -= 1
                n += 1
        self._tree_fat = None
        return n

    def _tree_replay(self, batch):
        """After a step: every proxy whose fat box the device replaced was taken out of the reference's tree and put back with
        that box (dynamictree.py:152-188), in synchronisation order = proxy row order (world.py:582-588)."""
        import numpy as np
        fat = batch.download(["f_fat"])["f_fat"][0]
        before = self._tree_fat
        if before is not None and before.shape == fat.shape:
            for j in np.nonzero((fat != before).any(axis=1))[0]:
                slot = self._tree_leaf.get(self._packed_keys[j])
                if slot is not None:
                    self._tree.reinsert_leaf(slot, fat[j])
        self._tree_fat = fat

    def _body_activation_changed(self, body):
        """`Body.active` setter (body.py:848-872): proxies created (fixture.py:250-261) or destroyed (263-267) in the tree replay;
        on a stepping world the device rows follow when the world is packed again (`_rebuild_batch`)."""
        for fixture in body.fixtures:
            if body._active:
                self._tree_create(body, fixture)
                self._activated.update((id(fixture), child) for child in range(fixture._shape.child_count))
            else:
                self._tree_destroy(fixture)
        self._host_dirty = True
        if self._stepped:
            self._structure_dirty = True

    def _fixture_refiltered(self, fixture):
        """`Fixture._refilter` (fixture.py:269-285) after a filter change: the fixture's contacts are flagged -- the next
        narrowphase pass that looks at them (one of the two bodies awake and not static, contact_manager.py:192-208) drops those
        that may no longer collide -- and its proxies go back into the move buffer, so that the pair pass at the end of the next
        step can create the contacts the new filter allows.  Host mutation of the device contact list: applied when the world is
        packed again before the next step (`_rebuild_batch`)."""
        self._host_dirty = True
        if not self._stepped or self._batch is None:
            return                              # no contacts yet, and every proxy is still in the move buffer
        D = self._batch.download(["c_fix", "w_count"])
        for k in range(int(D["w_count"][0, 2])):
            a, b = int(D["c_fix"][0, k, 0]), int(D["c_fix"][0, k, 1])
            if self._packed_fixtures[a] is fixture or self._packed_fixtures[b] is fixture:
                self._flagged_contacts.add((self._packed_keys[a], self._packed_keys[b]))
        for child in range(fixture._shape.child_count):
            self._touched_proxies.add((id(fixture), child))
        self._structure_dirty = True

    def _before_host_move(self, body):
        """The reference looks for new contacts at the end of EVERY transform setter (body.py:413), so the contacts of a body
        moved earlier exist, in that call's pair order, before the next body moves.  When proxies of ANOTHER body are still
        waiting for their pass, run it now: a narrowphase-only device step (dt = 0) with the new-fixture flag finds their
        contacts first; its narrowphase runs on the transforms as they are and is repeated unchanged by the next real step."""
        if self._moved_body is None or self._moved_body is body or not self._moved_proxies:
            return
        batch = self._ensure_batch()
        batch.inv_dt0 = self._inv_dt0
        batch.step(0.0, 1, 1, 1)
        batch.download_into_world(self)
        self._tree_fat = None

    def _body_moved_by_host(self, body):
        """`Body._transform_updated` (body.py:404-413): every proxy of the body is moved with no displacement
        (`Fixture._synchronize(xf, xf)`, fixture.py:287-298 -> `DynamicTree.move_proxy`, dynamictree.py:152-188: nothing happens
        while the old fat box still contains the new tight box, otherwise the leaf is re-inserted with the tight box grown by
        AABB_EXTENSION and the proxy enters the move buffer), then `find_new_contacts` runs.  The boxes are decided here, on the
        tree replay's leaf boxes (the device's fat boxes as of the last step), and uploaded before the next step with the world's
        new-fixture flag, which makes that step find the new contacts before its narrowphase -- the same contacts as the
        reference's immediate call; a second body moved before the next step first gets the first one's pass run
        (`_before_host_move`), so that every setter has its own pass, as there."""
        if not body._active:
            return
        tree = self._tree
        for fixture in body.fixtures:
            shape = fixture._shape
            for child in range(shape.child_count):
                key = (id(fixture), child)
                slot = self._tree_leaf.get(key)
                if slot is None:
                    continue
                tight = shape.compute_aabb(body._xf, child) if shape.child_count > 1 else shape.compute_aabb(body._xf)
                moved = tree.move_leaf(slot, tight, (0.0, 0.0), AABB_EXTENSION, 0.0)
                self._moved_proxies[key] = (tight, moved or self._moved_proxies.get(key, (None, False))[1])
        self._moved_body = body
        self._host_dirty = True

    def _upload_moved_proxies(self):
        import numpy as np
        batch = self._batch
        D = batch.download(["f_fat", "f_tight", "f_moved", "w_flags"])
        row = {k: j for j, k in enumerate(self._packed_keys)}
        for key, (tight, moved) in self._moved_proxies.items():
            j, slot = row.get(key), self._tree_leaf.get(key)
            if j is None or slot is None:
                continue
            D["f_tight"][0, j] = tight
            D["f_fat"][0, j] = self._tree._box(slot)
            if moved:
                D["f_moved"][0, j] = 1
        D["w_flags"] = D["w_flags"] | np.uint32(schema.WORLD_NEW_FIXTURE)
        batch.upload(D)
        self._moved_proxies = {}
        self._moved_body = None
        self._tree_fat = None

    def _ensure_batch(self):
        batch = self._ensure_batch_packed()
        if self._moved_proxies:
            self._upload_moved_proxies()
        return batch

    def _ensure_batch_packed(self):
        from .batch import BatchedWorld
        if self._batch is None:
            self._batch = BatchedWorld.from_worlds([self], real=self._real, device=self._device)
            self._host_dirty = False
            self._structure_dirty = False
            self._packed_fixtures = self._fixture_list()
            self._packed_keys = self._proxy_keys()
        elif self._structure_dirty:
            self._rebuild_batch()
        elif self._host_dirty:
            self._batch.upload_bodies_from_world(self)
            self._host_dirty = False
        return self._batch

    def _rebuild_batch(self):
        """Host mutation between steps (create_body / destroy_body / create_fixture / destroy_fixture after the
        first step, world.py:183-251, body.py:164-286): the world is packed again from the Python objects (which mirror the
        device state after every step) and what only the device knows is carried over by object identity -- the broadphase
        proxies of the surviving fixtures (fat / tight boxes, move buffer) and every contact whose two fixtures survive, in
        list order with its manifold, impulses and flags; a destroyed body's or fixture's contacts and proxies go with it
        (ContactManager.destroy, contact_manager.py:160-172), a new fixture gets a fresh proxy in the move buffer and raises
        the world's new-fixture flag (fixture.py:250-261, world.py:85)."""
        import numpy as np
        from .batch import BatchedWorld, world_caps
        old = self._batch
        D = old.download()
        old_index = {k: j for j, k in enumerate(self._packed_keys)}
        fixtures = self._fixture_list()
        keys = self._proxy_keys()
        caps = world_caps([self])
        caps = schema.Caps(max(caps.nb, 1), max(caps.nf, 1), max(caps.nc, old.caps.nc), nj=caps.nj)
        table = schema.ShapeTable()
        schema.add_shapes(table, self)
        A = schema.alloc(caps, 1, max(len(table), 1), "f64")
        schema.pack_world(self, A, 0, table)
        A["s_meta"], A["s_radius"], A["s_data"] = table.arrays("f64")
        remap = np.full(len(self._packed_fixtures) + 1, -1, np.int64)
        any_new = False
        for j2, f in enumerate(fixtures):
            j = old_index.get(keys[j2], -1)
            if j < 0:
                any_new = True
                continue
            remap[j] = j2
            if f._body._active and keys[j2] not in self._activated:
                for name in ("f_fat", "f_tight", "f_moved"):
                    A[name][0, j2] = D[name][0, j]
        self._activated = set()
        k2 = 0
        still_flagged = set()
        for k in range(int(D["w_count"][0, 2])):
            a2, b2 = int(remap[int(D["c_fix"][0, k, 0])]), int(remap[int(D["c_fix"][0, k, 1])])
            if a2 < 0 or b2 < 0:
                continue
            ba, bb = fixtures[a2]._body, fixtures[b2]._body
            if not (ba._active and bb._active):
                continue                       # Body.active = False took the body's contacts with it (body.py:868-872)
            if any(not j._collide_connected and j._body_a is not None and j.other_body(ba) is bb for j in ba.joints):
                continue                       # Body._should_collide (body.py:920-926), see add_joint
            ck = (keys[a2], keys[b2])
            if ck in self._flagged_contacts:   # contact_manager.py:192-208
                if (ba._awake and ba._type != Body.STATIC) or (bb._awake and bb._type != Body.STATIC):
                    if (ba._type != Body.DYNAMIC and bb._type != Body.DYNAMIC) or \
                            not self.default_contact_filter(fixtures[a2], fixtures[b2]):
                        continue
                else:
                    still_flagged.add(ck)      # both bodies rest: the pass skips the contact, the flag stays
            A["c_fix"][0, k2] = (a2, b2)
            for name in ("c_flags", "c_mat", "c_mtype", "c_mf", "c_pts", "c_imp", "c_id"):
                A[name][0, k2] = D[name][0, k]
            if fixtures[a2]._sensor or fixtures[b2]._sensor:       # Fixture.sensor may have been set since (contact.py:229-231)
                A["c_flags"][0, k2] |= schema.CONTACT_SENSOR
            else:
                A["c_flags"][0, k2] &= ~np.uint32(schema.CONTACT_SENSOR)
            k2 += 1
        A["w_count"][0, 2] = k2
        self._flagged_contacts = still_flagged
        for j2, key in enumerate(keys):
            if key in self._touched_proxies and fixtures[j2]._body._active:
                A["f_moved"][0, j2] = 1        # BroadPhase.touch_proxy (broadphase.py:70-72)
        self._touched_proxies = set()
        flags = int(A["w_flags"][0])
        if not any_new and not (int(D["w_flags"][0]) & schema.WORLD_NEW_FIXTURE):
            flags &= ~schema.WORLD_NEW_FIXTURE
        A["w_flags"][0] = flags
        batch = BatchedWorld(A, caps, real=self._real, device=self._device)
        batch.inv_dt0 = old.inv_dt0
        self._batch = batch
        self._packed_fixtures = fixtures
        self._packed_keys = keys
        self._host_dirty = False
        self._structure_dirty = bool(still_flagged)      # flagged contacts of resting pairs are looked at again before the next step
        self._tree_fat = None

    def step(self, dt, vel_iters, pos_iters):
        if self._locked:
            raise LockedError("step is not re-entrant")
        batch = self._ensure_batch()
        if self.controllers and dt > 0.0:
            # controllers act between the narrowphase and the island solver (world.py:338-342, 477-479): run the narrowphase
            # alone (a step with dt = 0: contacts updated, bodies woken by touching changes), mirror the awake flags back,
            # let the controllers apply their forces on the host as the reference does, upload them; the step below repeats the
            # narrowphase on unchanged transforms, which changes nothing
            batch.inv_dt0 = self._inv_dt0
            batch.step(0.0, vel_iters, pos_iters, 1)
            batch.download_into_world(self)
            step = TimeStep(dt, 1.0 / dt, self._inv_dt0 * dt, vel_iters, pos_iters, self._warm_starting)
            for controller in self.controllers:
                controller.step(step)
            batch = self._ensure_batch()
        self._locked = True
        try:
            if self._tree_fat is None:             # the device's boxes before the step (its fp32 boxes are not the host's fp64 ones)
                self._tree_fat = batch.download(["f_fat"])["f_fat"][0]
            batch.inv_dt0 = self._inv_dt0
            batch.step(dt, vel_iters, pos_iters, 1)
            if dt > 0.0:
                self._inv_dt0 = 1.0 / dt
            self._stepped = True
            # The device arrays have fixed capacities (the reference's lists have none): a step that ran out of contact,
            # pair or island-list rows dropped work and its result is NOT the reference's -- fail loudly, never mirror it
            # back (the read-back below synchronises anyway, so the check costs one small kernel).
            st = batch.status(raise_on_overflow=False)
            if st:
                names = [n for bit, n in ((schema.STATUS_CONTACT_OVERFLOW, "contacts"), (schema.STATUS_PAIR_OVERFLOW, "pairs"),
                                          (schema.STATUS_ISLAND_OVERFLOW, "island list"), (schema.STATUS_NAN, "NaN")) if st & bit]
                raise RuntimeError("World.step: device capacity exceeded (%s): contacts %d, pairs %d, island rows %d per world; "
                                   "the step's result is invalid" % (", ".join(names) or hex(st), batch.caps.nc, batch.caps.npair,
                                                                      batch.caps.nl))
            batch.download_into_world(self)
            self._tree_replay(batch)
            self._deliver_contact_events()
            if dt > 0.0:
                self._deliver_post_solve(batch)
            if not self._auto_clear_forces:
                # forces were cleared on the device only when auto_clear_forces is on
                pass
        finally:
            self._locked = False

    def clear_forces(self):
        for b in self.bodies:
            if b._force.x != 0.0 or b._force.y != 0.0 or b._torque != 0.0:
                self._host_dirty = True
            b._force = Vec2()
            b._torque = 0.0

    # -- queries (world.py:371-440) ----------------------------------------------------------------
    def _fixture_list(self):
        """The fixture of every device proxy row (a Loop fixture appears once per child edge, as in the reference's broadphase)."""
        return [f for f, c in schema.world_proxies(self)]

    def _proxy_keys(self):
        return [(id(f), c) for f, c in schema.world_proxies(self)]

    def query_aabb(self, aabb, callback=None):
        """Every fixture whose broadphase (fat) AABB overlaps `aabb` -- an object with lower_bound / upper_bound, or
        ((lx, ly), (ux, uy)).  A generator, or calls `callback(fixture)` until it returns a false value, as the reference
        (world.py:371-385).  Fixtures come in creation order (the reference: tree order; the set is the same)."""
        lo, hi = (aabb.lower_bound, aabb.upper_bound) if hasattr(aabb, "lower_bound") else aabb
        hits = self._query_aabb((lo[0], lo[1], hi[0], hi[1]))
        if callback is not None:
            for f in hits:
                if not callback(f):
                    break
            return None
        return iter(hits)

    def _query_aabb(self, box):
        batch = self._ensure_batch()
        fixtures = self._fixture_list()
        count, fix = batch.query_aabb([0], [box], max_hits=max(len(fixtures), 1))
        return [fixtures[j] for j in fix[0][:int(count[0])]]

    def ray_cast(self, point1, point2, callback=None):
        """Every fixture in the path of the ray, with the reference's protocol (world.py:387-440): yields (or passes to the
        callback) [fixture, point, normal, fraction]; a returned / assigned last entry of -1 ignores the fixture, 0 ends the
        cast, a fraction clips the ray there, 1 continues.  The shape ray casts run on the device; fixtures are visited in
        creation order (the reference: tree order, which Box2D leaves unspecified)."""
        p1, p2 = Vec2(*point1), Vec2(*point2)
        batch = self._ensure_batch()
        fixtures = self._fixture_list()
        count, fix, hit = batch.ray_cast([0], [(p1.x, p1.y, p2.x, p2.y)], max_hits=max(len(fixtures), 1))
        if int(count[0]) < 0:
            raise ValueError("||p2 - p1||^2 > 0.0")
        hits = [(fixtures[int(fix[0][k])], hit[0][k]) for k in range(int(count[0]))]

        def walk():
            max_fraction = 1.0
            for fixture, h in hits:
                fraction = float(h[0])
                if fraction > max_fraction:          # outside the clipped segment: the tree would have pruned it
                    continue
                ret = [fixture, Vec2(float(h[3]), float(h[4])), Vec2(float(h[1]), float(h[2])), fraction]
                if callback is not None:
                    value = callback(*ret)
                else:
                    yield ret
                    value = ret[-1]
                if value == 0.0:
                    return
                if value > 0.0:
                    max_fraction = value

        if callback is not None:
            for _ in walk():
                pass
            return None
        return walk()

    def _listeners_changed(self):
        # a listener attached to a world that is already stepping hears about changes from now on
        if self._batch is not None and self._stepped and self._prev_contacts is None:
            self._prev_contacts = self._contact_views()

    def _deliver_contact_events(self):
        """begin_contact / end_contact for the step that just ran.  Within one step the reference touches every contact that
        existed before it at most once in ContactManager.collide, in list order: destroyed (end_contact if it was touching,
        contact_manager.py:167-168) or updated (begin / end on a touching change, contact.py:263-266); contacts created by the
        step are not touching yet.  So the events are the ordered difference of the two contact lists.  (With continuous
        collision on, a TOI sub-step can change a flag again inside the step: only the net change is reported.)"""
        cm = self.contact_manager
        if cm.begin_contact is None and cm.end_contact is None:
            self._prev_contacts = None
            return
        cur = self._contact_views()
        prev = self._prev_contacts
        self._prev_contacts = cur
        if prev is None:
            prev = []
        now = {(id(c.fixture_a), id(c.fixture_b)): c for c in cur}
        for old in prev:
            c = now.get((id(old.fixture_a), id(old.fixture_b)))
            if c is None:
                if old.touching and cm.end_contact is not None:
                    cm.end_contact(old)
            elif c.touching and not old.touching:
                if cm.begin_contact is not None:
                    cm.begin_contact(c)
            elif old.touching and not c.touching:
                if cm.end_contact is not None:
                    cm.end_contact(c)
        seen = set((id(c.fixture_a), id(c.fixture_b)) for c in prev)
        for c in cur:                                   # (a contact cannot be created touching; kept for robustness)
            if (id(c.fixture_a), id(c.fixture_b)) not in seen and c.touching and cm.begin_contact is not None:
                cm.begin_contact(c)

    def _deliver_post_solve(self, batch):
        """island.py:245-252 for every island the step solved, in island order (world.py:574): the step's end only appends
        contacts, so the island lists' contact rows still name the same contacts; the impulses are the ones the solver stored
        (contact.py:801-810).  Not reproduced: the reports of the TOI sub-islands of a world with continuous collision
        (island.py:243), and the shortened impulse list of a two-point contact whose block matrix was ill-conditioned
        (contact.py:507-510: both stored impulses are handed out here)."""
        fcn = self.contact_manager._post_solve
        if fcn is None:
            return
        D = batch.download(["w_island", "l_contact"])
        n = int(D["w_island"][0, 2])
        if n == 0:
            return
        views = self._contact_views()
        for k in range(n):
            c = views[int(D["l_contact"][0, k])]
            fcn(c, [(ni, ti) for _, ni, ti in c.points])

    def _contact_views(self):
        if self._batch is None:
            return []
        return self._batch.contact_views(self)

    @staticmethod
    def default_contact_filter(filter_a, filter_b):
        # world.py:958-970; the device evaluates the same rule in the pair-filter kernel
        if filter_a._group_index == filter_b._group_index and filter_a._group_index != 0:
            return filter_a._group_index > 0
        return ((filter_a._mask_bits & filter_b._category_bits) != 0 and
                (filter_a._category_bits & filter_b._mask_bits) != 0)
