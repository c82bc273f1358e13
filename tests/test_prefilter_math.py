"""The inequality behind the 8-bit prefilter of the sweep kernels (particlesmc_b200/csrc/common.cuh: pack8, neg_thr8;
chains_spec.cuh: the filter sphere of a Displacement and of a MoleculeFlip), restated in numpy and checked on random
inputs: a candidate within the cutoff of the old or the new position (Displacement), or of either site (MoleculeFlip),
always passes the byte test around the fixed-point midpoint -- the prefilter may only ADD candidates to what the
reference evaluates (src/atoms.jl:66-88 visits every neighbour and applies r2 <= rcut2 itself), never drop one.
CPU only: this is the margin argument, not the kernel."""
import numpy as np


def fixed32(x, L):
    return np.floor(x * (4294967296.0 / L)).astype(np.uint64) & 0xFFFFFFFF


def byte_test(centre_u32, cand_u32, r_units):
    """VABSDIFF4 + IDP.4A on the top bytes: sum of squared wrapped byte differences <= (r + sqrt 3)^2 + 1."""
    cb = (centre_u32 >> 24).astype(np.int64)
    kb = (cand_u32 >> 24).astype(np.int64)
    d = np.abs(cb - kb)
    d = np.where(d >= 128, d - 256, d)  # the signed-byte reading of IDP.4A
    r2q = (d * d).sum(axis=-1)
    t = r_units + 1.7320526
    thr = np.minimum(np.floor(t * t + 1.0), 60000.0)
    return r2q <= thr


def midpoint(u_a, u_b):
    e = ((u_b.astype(np.int64) - u_a.astype(np.int64) + (1 << 31)) % (1 << 32)) - (1 << 31)  # wrapping int32 difference
    return (u_a.astype(np.int64) + (e >> 1)) % (1 << 32), e


def mi_dist(a, b, L):
    d = a - b
    d -= np.round(d / L) * L
    return np.sqrt((d * d).sum(axis=-1))


def candidates_near(rng, centre, rc, L, n):
    """Points at distance <= rc of `centre` (biased towards the shell, where the margin matters), folded into the box."""
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    r = rc * (1.0 - rng.random(n) ** 4 * 0.2)
    return np.mod(centre + v * r[:, None], L)


def test_displacement_sphere_never_drops_an_in_range_candidate():
    rng = np.random.default_rng(1)
    n = 200_000
    for L, rc in [(9.41036, 2.5), (4.2, 2.0), (32.8962, 3.5)]:
        xo = rng.random((n, 3)) * L
        delta = rng.normal(0, 0.08, (n, 3))
        xn = np.mod(xo + delta, L)
        uo = fixed32(xo, L)
        di = np.rint(delta * (4294967296.0 / L)).astype(np.int64)  # ri[] of the proposal record
        um = (uo.astype(np.int64) + (di >> 1)) % (1 << 32)
        hd = 0.5 * np.linalg.norm(delta, axis=1)
        r_units = (rc + hd) * 256.0 / L
        for centre in (xo, xn):
            k = candidates_near(rng, centre, rc, L, n)
            assert np.all(mi_dist(k, centre, L) <= rc + 1e-12)
            assert np.all(byte_test(um, fixed32(k, L), r_units))


def test_flip_sphere_around_the_bond_midpoint_holds_both_sites_spheres():
    rng = np.random.default_rng(2)
    n = 200_000
    L, rc = 13.572, 1.2347  # the trimer fixture: WCA cutoff 2^(1/6) * 1.1, FENE bonds up to r0 = 1.575
    xi = rng.random((n, 3)) * L
    b = rng.normal(size=(n, 3))
    b *= (rng.random(n) * 1.6 / np.linalg.norm(b, axis=1))[:, None]
    xj = np.mod(xi + b, L)  # often across a periodic boundary
    ui, uj = fixed32(xi, L), fixed32(xj, L)
    um, e = midpoint(ui, uj)
    hd = 0.5 * np.sqrt((e.astype(np.float64) ** 2).sum(axis=1)) * 2.0 ** -24
    r_units = rc * 256.0 / L + hd + 2.0 ** -20
    for centre in (xi, xj):
        k = candidates_near(rng, centre, rc, L, n)
        assert np.all(byte_test(um, fixed32(k, L), r_units))
    # and the sphere is not vacuous: far candidates are rejected
    far = np.mod(xi + L / 2, L)
    assert not np.any(byte_test(um, fixed32(far, L), r_units))
