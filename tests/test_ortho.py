"""Row (f)1 of the scope table: batched orthogonal / nullspace (Orthogonal.h:73-178).
CPU: oracle.c against the reference-generated fixtures, against the compiled reference, and the
reference's own regression property (regression/linear_solve.cpp:121-150: every null basis has a
zero dot product with every input basis, abs 1e-5).  GPU: the CUDA path against both."""
import numpy as np
import pytest

from conftest import GOLDEN_DIR, bits_equal, first_diff
import os


def golden(tag):
    return np.load(os.path.join(GOLDEN_DIR, f"ortho_{tag}.npz"))


def cases(g):
    orth = sorted(int(k.split("_")[2]) for k in g.files if k.startswith("orth_v_"))
    ns = sorted(tuple(int(x) for x in k.split("_")[2:]) for k in g.files if k.startswith("ns_b_"))
    return orth, ns


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_port_matches_golden(port, tag):
    g = golden(tag)
    orth, ns = cases(g)
    with np.errstate(all="ignore"):
        for n in orth:
            got = port.orthogonal(g[f"orth_v_{n}"], n)
            assert bits_equal(got, g[f"orth_o_{n}"]), f"orthogonal n={n}: {first_diff(got, g[f'orth_o_{n}'])}"
        for n, nb in ns:
            x, ok = port.nullspace(g[f"ns_b_{n}_{nb}"], n, nb)
            assert bits_equal(x, g[f"ns_x_{n}_{nb}"]), f"nullspace N={n} nb={nb}: {first_diff(x, g[f'ns_x_{n}_{nb}'])}"
            assert np.array_equal(ok, g[f"ns_ok_{n}_{nb}"])
    # zero bases are detected; a dependent LAST basis is not (the last diagonal is never tested, LUDecomp.h:203)
    assert (g["orth_o_5"][0] == 0).all() and g["ns_ok_5_2"][0] == 0 and g["ns_ok_5_2"][1] == 1 and g["ns_ok_5_2"][2] == 1


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_port_equals_reference(port, ref, dt):
    rng = np.random.default_rng(11)
    with np.errstate(all="ignore"):
        for n in range(2, 17):
            v = rng.standard_normal((300, (n - 1) * n)).astype(dt)
            v[::31] = 0
            assert bits_equal(port.orthogonal(v, n), ref.orthogonal(v, n)), n
            for nb in range(1, n if n >= 3 else 0):
                bs = rng.standard_normal((200, nb * n)).astype(dt)
                bs[::29] = 0
                a, b = port.nullspace(bs, n, nb), ref.nullspace(bs, n, nb)
                assert bits_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (n, nb)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 10])
def test_verify_nullspace_property(port, n):
    """regression/linear_solve.cpp:121-150, double, N(0,1) entries, 256 trials per N."""
    rng = np.random.default_rng(16512485420001724907 % (2**32) + n)
    for _ in range(8):
        nb = int(rng.integers(1, n)) if n > 2 else 1
        bs = rng.standard_normal((32, nb * n))
        if n == 2:
            x = port.orthogonal(bs, 2)
        else:
            x, ok = port.nullspace(bs, n, nb)
            assert ok.all()
        dots = np.einsum("sik,sjk->sij", bs.reshape(32, nb, n), x.reshape(32, n - nb, n))
        assert np.abs(dots).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("soa", [False, True])
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_gpu_matches_golden_and_oracle(port, orc, tag, soa):
    torch = pytest.importorskip("torch")
    import geomc_b200 as gm
    gm.lib()

    def up(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        return gm.BatchSoA.from_aos(t) if soa else t

    def down(x):
        return (x.to_aos() if hasattr(x, "to_aos") else x).cpu().numpy()

    g = golden(tag)
    orth, ns = cases(g)
    dt = np.float32 if tag == "f32" else np.float64
    rng = np.random.default_rng(5)
    with np.errstate(all="ignore"):
        for n in orth:
            got = down(gm.orthogonal(up(g[f"orth_v_{n}"]), n))
            assert bits_equal(got, g[f"orth_o_{n}"]), f"orthogonal n={n} soa={soa}: {first_diff(got, g[f'orth_o_{n}'])}"
        for n, nb in ns:
            x, ok = gm.nullspace(up(g[f"ns_b_{n}_{nb}"]), n, nb)
            assert bits_equal(down(x), g[f"ns_x_{n}_{nb}"]), f"nullspace N={n} nb={nb} soa={soa}"
            assert np.array_equal(down(ok), g[f"ns_ok_{n}_{nb}"])
        for n in range(2, 17):                                # ragged random batches against oracle.c
            v = rng.standard_normal((1000 + 37, (n - 1) * n)).astype(dt)
            v[::97] = 0
            assert bits_equal(down(gm.orthogonal(up(v), n)), port.orthogonal(v, n)), f"random orthogonal n={n}"
            if n >= 3:
                nb = max(1, n // 2)
                bs = rng.standard_normal((333, nb * n)).astype(dt)
                bs[::29] = 0
                x, ok = gm.nullspace(up(bs), n, nb)
                xe, oke = port.nullspace(bs, n, nb)
                assert bits_equal(down(x), xe) and np.array_equal(down(ok), oke), f"random nullspace N={n} nb={nb}"
        # every basis count for the sizes served by the register kernel (k_sub_plu MODE_RECT: float N <= 10, double N <= 7) and
        # its neighbours, against the compiled reference: dependent vectors, zero vectors, ties, huge / tiny scalings
        for n in range(5, 12):
            for nb in range(1, n):
                bs = rng.standard_normal((257, nb * n)).astype(dt)
                bs[::29] = 0
                bs[5::31, :n] = 1.0                                  # ties in the first vector
                if nb >= 2:
                    bs[7::37, n:2 * n] = bs[7::37, :n]               # two equal vectors: degenerate
                bs[11::41] *= dt(1e20 if dt == np.float32 else 1e150)
                bs[13::43] *= dt(1e-20 if dt == np.float32 else 1e-150)
                x, ok = gm.nullspace(up(bs), n, nb)
                xe, oke = orc.nullspace(bs, n, nb)
                assert bits_equal(down(x), xe), f"nullspace vs {orc.kind} N={n} nb={nb} soa={soa}: {first_diff(down(x), xe)}"
                assert np.array_equal(down(ok), oke), f"nullspace ok N={n} nb={nb}"
            v = rng.standard_normal((513, (n - 1) * n)).astype(dt)
            v[3::17, n:2 * n] = v[3::17, :n]
            v[11::41] *= dt(1e20 if dt == np.float32 else 1e150)
            oe = orc.orthogonal(v, n)
            got = down(gm.orthogonal(up(v), n))
            assert bits_equal(got, oe), f"orthogonal vs {orc.kind} n={n} soa={soa}: {first_diff(got, oe)}"
