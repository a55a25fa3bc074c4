"""CPU tests: pin the kinematic-chain restatement (oracle/oracle_chain.cpp, chain metric of oracle/oracle_knn.cpp) on the
golden vectors produced by the UNMODIFIED reference (demos/KinematicChain.h + DiscreteMotionValidator.cpp compiled in
oracle/_ref/libref_chain.so) and, where that build exists, live; plus the engine's host-side horn environment."""
import os

import numpy as np
import pytest

import ompl_b200
from oracle import chain, port

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain.npz"))
N = int(Z["n"])


def make(cls):
    return cls(N, 1.0 / N, Z["env"])


def test_horn_environment_port_and_engine_match_reference():
    np.testing.assert_array_equal(chain.PortChain.horn(N, np.log(N) / N), Z["env"])
    np.testing.assert_array_equal(ompl_b200.horn_environment(N, np.log(N) / N), Z["env"])   # b200mp_chain_horn_environment (host)
    assert Z["env"].shape == (2 * (N - 1), 4)


def test_port_matches_golden():
    P = make(chain.PortChain)
    assert P.longest_valid_segment == float(Z["lvs"])
    np.testing.assert_array_equal(P.validate_states(Z["states"]), Z["states_valid"])
    np.testing.assert_array_equal(P.distance(Z["a"], Z["b"]), Z["dist"])
    np.testing.assert_array_equal(P.interpolate(Z["a"], Z["b"], Z["t"]), Z["interp"])
    v0, _, nd = P.check_motions(Z["a"], Z["b"], 0)
    v1, last, _ = P.check_motions(Z["a"], Z["b"], 1)
    np.testing.assert_array_equal(v0, Z["motions_valid"])
    np.testing.assert_array_equal(v1, Z["motions_valid_form1"])
    np.testing.assert_array_equal(nd, Z["nd"])
    np.testing.assert_array_equal(last, Z["last_t"])
    np.testing.assert_array_equal(v0, v1)                     # both forms of checkMotion agree on the boolean
    assert 0.1 < v0.mean() < 0.9 and 0.05 < Z["states_valid"].mean() < 0.95


def test_chain_metric_knn_of_the_port_uses_the_reference_distance():
    db, qs = Z["states"][:400], Z["states"][400:420]
    idx, dist = port.knn(db, qs, 8, metric=1, param=1.0 / N)
    P = make(chain.PortChain)
    for r in range(len(qs)):
        d = P.distance(db, np.repeat(qs[r][None], len(db), 0))     # distFun(element, query)
        order = np.lexsort((np.arange(len(db)), d))[:8]
        np.testing.assert_array_equal(idx[r], order)
        np.testing.assert_array_equal(dist[r], d[order])


@pytest.mark.skipif(not chain.ref_available(), reason="oracle/_ref/libref_chain.so not built")
def test_port_matches_reference_live():
    R, P = make(chain.RefChain), make(chain.PortChain)
    rng = np.random.default_rng(1)
    q = rng.uniform(-np.pi, np.pi, (4000, N)) * rng.uniform(0.01, 1.0, (4000, 1))
    np.testing.assert_array_equal(R.validate_states(q), P.validate_states(q))
    a, b = q[:2000], q[2000:]
    np.testing.assert_array_equal(R.distance(a, b), P.distance(a, b))
    np.testing.assert_array_equal(R.interpolate(a, b, 0.731), P.interpolate(a, b, 0.731))
    for form in (0, 1):
        for x, y in zip(R.check_motions(a[:600], b[:600], form), P.check_motions(a[:600], b[:600], form)):
            np.testing.assert_array_equal(x, y)
    # other chain sizes and an empty environment
    for n in (3, 8, 31):
        R2, P2 = chain.RefChain(n, 1.0 / n, np.zeros((0, 4))), chain.PortChain(n, 1.0 / n, np.zeros((0, 4)))
        q2 = rng.uniform(-np.pi, np.pi, (1500, n))
        np.testing.assert_array_equal(R2.validate_states(q2), P2.validate_states(q2))
        np.testing.assert_array_equal(R2.distance(q2[:700], q2[700:1400]), P2.distance(q2[:700], q2[700:1400]))
