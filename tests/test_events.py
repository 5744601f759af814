"""Contact events: begin_contact / end_contact listeners on the facade against the LIVE reference's callback
sequence (tests/golden_queries/events.npz, oracle/gen_events_golden.py): same events, same order, same step."""
import os

import numpy as np
import pytest

import gen_events_golden as ge


def test_pre_solve_fails_loudly_post_solve_is_accepted():
    import pypybox2d_b200 as b2
    w = b2.World()
    w.contact_manager.begin_contact = lambda c: None
    assert w.contact_manager.contact_listener[0] is not None
    with pytest.raises(NotImplementedError):
        w.contact_manager.pre_solve = lambda c, m: None
    w.contact_manager.post_solve = fcn = lambda c, i: None       # delivered after the step (island.py:245-252)
    assert w.contact_manager.post_solve is fcn and w.contact_manager.contact_listener[3] is fcn


@pytest.mark.gpu
def test_post_solve_calls_match_the_reference():
    """tests/golden_queries/postsolve.npz (oracle/gen_postsolve_golden.py): the live reference's post_solve calls of an 8-row
    pyramid landing and settling -- same contacts in the same (island) order at every step, same impulses."""
    import gen_postsolve_golden as gp
    import pypybox2d_b200 as b2
    from pypybox2d_b200 import scenes
    want = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_queries", "postsolve.npz"))["rows"]
    got = gp.run(b2, scenes.pyramid(b2, world=1, rows=8, real="f64"))
    assert got.shape == want.shape, "number of post_solve calls: %d against %d" % (len(got), len(want))
    assert np.array_equal(got[:, :4], want[:, :4]), "step, fixtures and point count of every call, in call order"
    scale = np.abs(want[:, 4:]).max()
    worst = float(np.abs(got[:, 4:] - want[:, 4:]).max())
    assert worst <= 1e-9 * scale, "impulses: %.2e of %.2e" % (worst, scale)
    print("post_solve: %d calls, worst impulse difference %.2e (largest impulse %.2f)" % (len(got), worst, scale))


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(ge.CASES)))
def test_contact_events_match_the_reference(k):
    import pypybox2d_b200 as b2
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_queries", "events.npz"))
    scene, world, kw, dt, vi, pi, steps = ge.CASES[k]
    got = ge.run(b2, scene, world, dict(kw, real="f64"), dt, vi, pi, steps)
    want = G["e%d" % k]
    if scene == "bullets":
        # continuous collision on: a TOI sub-step can flip a touching flag a second time inside a step; the facade reports the
        # net change per step.  Compare the per-step net effect on every contact.
        def net(ev):
            out = {}
            for s, kind, fa, fb in ev.tolist():
                out[(s, fa, fb)] = out.get((s, fa, fb), 0) + (1 if kind == 0 else -1)
            return {key: v for key, v in out.items() if v != 0}
        assert net(got) == net(want)
    else:
        assert got.shape == want.shape and np.array_equal(got, want)
