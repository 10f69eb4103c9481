"""Property tests (hypothesis) that pin the CPU oracle to numpy statements of the reference's formulas on random
inputs: room rasteriser (simulations.py:536-576), gradient-to-velocity map (optimals.py:168-186), sampler index
semantics incl. its IndexError cases (optimals.py:212-250), nearest-wall argmin with first-index ties
(pedestrians.py:311-313) and the density splat (simulations.py:469-487).  The goldens fix a handful of rooms; these
cover the input space around them.  (The CUDA path is compared with the same oracle on the GPU box.)"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import cpu_oracle as co

SET = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


def _grid(L, H, step=0.05):
    Nx, Ny = int(L // step + 1), int(H // step + 1)      # simulations.py:63-64
    return np.linspace(0, L, Nx), np.linspace(0, H, Ny)


rect = st.tuples(st.floats(0.0, 6.0), st.floats(0.0, 4.0), st.floats(0.05, 3.0), st.floats(0.05, 3.0))
disc = st.tuples(st.floats(0.0, 6.0), st.floats(0.0, 4.0), st.floats(0.05, 1.0))


@settings(**SET)
@given(st.floats(3.0, 6.0), st.floats(2.0, 4.0), st.lists(rect, max_size=3), st.lists(rect, max_size=2),
       st.lists(disc, max_size=3), st.lists(rect, min_size=1, max_size=2))
def test_rasteriser_equals_numpy_masks(L, H, walls, holes, cyls, targets):
    X, Y = _grid(L, H)
    XX, YY = np.meshgrid(X, Y)
    inside = lambda r: (np.abs(XX - r[0]) < r[2] / 2) & (np.abs(YY - r[1]) < r[3] / 2)   # strict, on linspace nodes
    V = np.zeros_like(XX)
    for w in walls:
        V += np.where(inside(w), -1.0, 0.0)
    for h in holes:
        V[inside(h)] = 0
    for c in cyls:
        V += np.where(np.sqrt((XX - c[0]) ** 2 + (YY - c[1]) ** 2) < c[2], -1.0, 0.0)
    V[:, 0] = V[:, -1] = V[0, :] = V[-1, :] = -1
    for t in targets:
        V[inside(t)] = 1
    got = co.create_potential(X, Y, [list(w) for w in walls], [list(h) for h in holes], [list(c) for c in cyls],
                              [list(t) for t in targets])
    assert np.array_equal(got, V)


@settings(**SET)
@given(st.integers(4, 30), st.integers(4, 30), st.integers(0, 2 ** 31 - 1), st.floats(0.5, 10.0))
def test_vels_equals_numpy_statement(ny, nx, seed, mu):
    rng = np.random.RandomState(seed)
    lim, dx = 10e-3, 0.05
    old = np.seterr(all="ignore")   # the reference's formula divides by zero / multiplies inf by 0 at clamp edges
    phi = np.exp(rng.normal(0, 2.0, (ny, nx)))
    phi[rng.uniform(size=phi.shape) < 0.1] = 1e-3 * rng.uniform(0.1, 5)     # some cells below the clamp
    p = phi * (phi > lim) + lim * (phi < lim)
    gx = (p[1:-1, 2:] - p[1:-1, :-2]) / (2 * dx)
    gy = (p[2:, 1:-1] - p[:-2, 1:-1]) / (2 * dx)
    pc = p[1:-1, 1:-1] * (p[1:-1, 1:-1] > lim) + lim * (p[1:-1, 1:-1] < lim)
    vx, vy = gx / (mu * pc), gy / (mu * pc)
    n = np.sqrt(vx ** 2 + vy ** 2)
    den = n * (n > lim) + (n < lim)
    ex, ey = (vx * (n > lim)) / den, (vy * (n > lim)) / den
    ox, oy = co.vels(phi.ravel(), ny, nx, dx, dx, mu, lim)
    ok = np.abs(n - lim) > 1e-12                                            # n == lim exactly is 0/0 in the reference
    np.seterr(**old)
    np.testing.assert_allclose(ox[ok], ex[ok], rtol=1e-14, atol=0, equal_nan=True)
    np.testing.assert_allclose(oy[ok], ey[ok], rtol=1e-14, atol=0, equal_nan=True)


def _sampler_spec(vx, vy, nt_opt, L, H, dx, x, y, t):
    """optimals.py:212-250 with numpy's own fancy indexing; returns None where numpy raises IndexError"""
    Ny, Nx = vx.shape[1] + 2, vx.shape[2] + 2
    if t >= nt_opt - 1:
        return (0.0, 0.0)
    if x < L - dx:
        j = int(x // dx)
        if x > dx:
            j = [j, j + 1]
    else:
        j = Nx - 3
    if y < H - dx:
        i = int(y // dx)
        if y > dx:
            i = [i, i + 1]
    else:
        i = Ny - 3
    try:
        return (float(np.mean(vx[t][i, j])), float(np.mean(vy[t][i, j])))
    except IndexError:
        return None


@settings(**SET)
@given(st.floats(2.0, 5.0), st.floats(1.5, 4.0), st.integers(0, 2 ** 31 - 1))
def test_sampler_index_semantics(L, H, seed):
    cfg = {"dt": 0.02, "grid_step": 0.05, "relaxation": 0.1, "v_max": 2.0, "repulsion_cutoff": 4.0, "b_min": 0.2,
           "tau_a": 0.5, "b_max": 0.4, "eta": 0.2, "eta_walls": 0.1, "hjb_params": {"sigma": 0.2}}
    X, Y = _grid(L, H)
    Nx, Ny = len(X), len(Y)
    rng = np.random.RandomState(seed)
    nsl = 3
    vx, vy = rng.normal(size=(nsl, Ny - 2, Nx - 2)), rng.normal(size=(nsl, Ny - 2, Nx - 2))
    P = co.gcfm_params(cfg, L, H, Ny, Nx)
    key = co.KeyData(np.zeros((Ny, Nx)), vx, vy, nsl + 1, [[L, H / 2, 0.5, 0.5]])
    # incl. positions outside the room: numpy wraps the negative indices of x < 0 / y < 0 silently (App. C #7)
    xs = np.concatenate([rng.uniform(0, L, 40), [0.0, 0.05, 0.049999, 0.050001, L - 0.05, L - 0.1, L - 0.0999, L],
                         [-0.01, -0.05, -0.3, 1.0, -1e-9, -L, -L - 1.0, L + 0.3]])
    ys = np.concatenate([rng.uniform(0, H, 40), [0.0, 0.05, H - 0.05, H - 0.1, H - 0.0999, H, 0.1, 0.15000000000000002],
                         [0.5, -0.07, -0.02, -0.2, 0.3, 0.4, 0.5, -H - 0.5]])
    for x, y in zip(xs, ys):
        for t in (0, nsl - 1, nsl, nsl + 3):
            want = _sampler_spec(vx, vy, nsl + 1, L, H, 0.05, float(x), float(y), t)
            ox, oy, bad = co.choose_velocity(P, key, float(x), float(y), t)
            if want is None:
                assert bad, (x, y, t)
            else:
                assert not bad and (ox, oy) == want, (x, y, t, (ox, oy), want)


@settings(**SET)
@given(st.floats(2.0, 5.0), st.floats(1.5, 4.0), st.integers(0, 2 ** 31 - 1))
def test_wall_argmin_first_index_ties(L, H, seed):
    cfg = {"dt": 0.02, "grid_step": 0.05, "relaxation": 0.1, "v_max": 2.0, "repulsion_cutoff": 4.0, "b_min": 0.2,
           "tau_a": 0.5, "b_max": 0.4, "eta": 0.2, "eta_walls": 0.1, "hjb_params": {"sigma": 0.2}}
    X, Y = _grid(L, H)
    XX, YY = np.meshgrid(X, Y)
    rng = np.random.RandomState(seed)
    V = np.zeros_like(XX)
    V[:, 0] = V[:, -1] = V[0, :] = V[-1, :] = -100
    for _ in range(rng.randint(0, 4)):
        cx, cy, r = rng.uniform(0.3, L - 0.3), rng.uniform(0.3, H - 0.3), rng.uniform(0.1, 0.5)
        V[np.sqrt((XX - cx) ** 2 + (YY - cy) ** 2) < r] = -100
    V[len(Y) // 2 - 2:len(Y) // 2 + 2, -1] = 1
    P = co.gcfm_params(cfg, L, H, len(Y), len(X))
    pts = np.column_stack([rng.uniform(0.2, L - 0.2, 12), rng.uniform(0.2, H - 0.2, 12)])
    pts[:4] = np.round(pts[:4] / 0.05) * (L / (len(X) - 1))          # on / between nodes: distance ties
    for p in pts:
        if V[np.argmin(np.abs(Y - p[1])), np.argmin(np.abs(X - p[0]))] < 0:
            continue                                                  # standing on a wall node is R = 0 (NaN) in the reference
        want = int(np.argmin(np.sqrt((XX - p[0]) ** 2 + (YY - p[1]) ** 2) + V * 10e3))   # pedestrians.py:311-313
        _, _, ind = co.wall_force(P, X, Y, V, p, (0.3, -0.2), 1.3)
        assert ind == want, (p, ind, want)


@settings(**SET)
@given(st.floats(2.0, 4.0), st.floats(1.5, 3.0), st.integers(0, 2 ** 31 - 1), st.floats(0.2, 0.8))
def test_density_equals_numpy_statement(L, H, seed, sigma):
    X, Y = _grid(L, H)
    XX, YY = np.meshgrid(X, Y)
    rng = np.random.RandomState(seed)
    n = rng.randint(0, 6)
    x, y = rng.uniform(0, L, n), rng.uniform(0, H, n)
    status = (rng.uniform(size=n) < 0.7).astype(np.uint8)
    Vg = np.where(rng.uniform(size=XX.shape) < 0.1, -3.0, 0.0)
    d = np.zeros_like(XX)
    for q in range(n):
        if status[q]:
            d += np.exp(-((XX - x[q]) ** 2 + (YY - y[q]) ** 2) / (2 * sigma ** 2)) / np.sqrt(4 * np.pi ** 2 * sigma ** 2)
    d[Vg < 0] = 0
    got = co.density(X, Y, Vg, x, y, status, sigma)
    np.testing.assert_allclose(got, d, rtol=2e-15, atol=1e-300)      # own exp: <= 1-2 ulp from numpy's


# ---- round 2: the conservative pruning rules of the GCFM candidate search and of the nearest-wall search, restated in
# numpy (csrc/oc_gcfm.cu: CandSearch::gather, WallSearch::tile_lb).  The kernels are compared with the sequential sweep
# bit for bit on the GPU box; these check the geometry behind the rules on the whole input space.
def _fov_cull(vx, vy, ox, oy, margin, cos_fov):
    """the kernel's predicate: a candidate whose OLD offset from agent i is (ox, oy) cannot enter i's field of view
    wherever it moves within rho = margin / sqrt(2)"""
    rho_s = margin * 0.70710678118654757 * (1.0 + 1e-6) + 1e-12
    sin_fov = np.sqrt(max(1.0 - cos_fov * cos_fov, 0.0))
    ni = np.sqrt(vy * vy + vx * vx)
    if ni == 0.0:
        return True
    d = np.sqrt(ox * ox + oy * oy)
    if d == 0.0:
        return False
    sn = rho_s / d
    if not sn < 0.999 * sin_fov:
        return False
    cs = np.sqrt(1.0 - sn * sn)
    cos_lim = cos_fov * cs - sin_fov * sn
    return (vx * ox + vy * oy) / (ni * d) < cos_lim - 1e-9


@settings(max_examples=300, deadline=None)
@given(st.floats(-2.0, 2.0), st.floats(-2.0, 2.0), st.floats(0.05, 5.0), st.floats(0.0, 2 * np.pi),
       st.floats(0.0, 1.0), st.floats(0.0, 2 * np.pi), st.sampled_from([0.06, 0.25, 1.0]))
def test_field_of_view_culling_is_conservative(vx, vy, d, ang, frac, dang, margin):
    """whenever the rule drops a candidate, the reference's k (pedestrians.py:259-262) is exactly 0 for EVERY position
    the candidate can reach in the step (displacement <= margin/2 per axis, i.e. <= margin/sqrt(2)), so its force is
    -+0.0 and skipping it changes no bit of the agent's sum"""
    cos_fov = np.cos(0.7 * np.pi)
    ox, oy = d * np.cos(ang), d * np.sin(ang)
    if not _fov_cull(vx, vy, ox, oy, margin, cos_fov):
        return
    rho = margin / np.sqrt(2.0)
    for f in (frac, 1.0):                       # a random displacement and one on the boundary of the disc
        Rx, Ry = ox + f * rho * np.cos(dang), oy + f * rho * np.sin(dang)
        nR = np.sqrt(Ry * Ry + Rx * Rx)
        ni = np.sqrt(vy * vy + vx * vx)
        k = 0.0
        if ni > 0:
            ex, ey = Rx / nR, Ry / nR
            k = np.maximum((vx * ex + vy * ey) / ni - cos_fov, 0) / (1 - cos_fov)
        assert k == 0.0


def test_field_of_view_culling_is_not_vacuous():
    """the rule drops about a quarter of uniformly placed candidates at the default margin (0.7 pi half-angle)"""
    rng = np.random.RandomState(0)
    cos_fov = np.cos(0.7 * np.pi)
    n = hit = 0
    for _ in range(4000):
        th, r = rng.uniform(0, 2 * np.pi), 4.2 * np.sqrt(rng.uniform(0.01, 1))
        hit += _fov_cull(1.2, 0.3, r * np.cos(th), r * np.sin(th), 0.25, cos_fov)
        n += 1
    assert 0.2 < hit / n < 0.3


@settings(**SET)
@given(st.floats(3.0, 9.0), st.floats(2.0, 6.0), st.integers(0, 10 ** 6))
def test_wall_tile_lower_bound_holds_in_floating_point(L, H, seed):
    """WallSearch::tile_lb: for every node of a 16 x 16 tile the COMPUTED distance sqrt(ddx*ddx + ddy*ddy) is >= the
    bound sqrt(gx*gx + gy*gy) with gx = max(X[ix0] - x, x - X[ix1], 0) evaluated with the same operations (rounding is
    monotonic), so a tile whose bound exceeds the best key cannot hold the argmin nor a tie"""
    X, Y = _grid(L, H)
    rng = np.random.RandomState(seed)
    WT = 16
    for _ in range(20):
        x, y = rng.uniform(-0.5, L + 0.5), rng.uniform(-0.5, H + 0.5)
        tx, ty = rng.randint(0, (len(X) + WT - 1) // WT), rng.randint(0, (len(Y) + WT - 1) // WT)
        ix0, iy0 = tx * WT, ty * WT
        ix1, iy1 = min(ix0 + WT - 1, len(X) - 1), min(iy0 + WT - 1, len(Y) - 1)
        gx = max(max(X[ix0] - x, x - X[ix1]), 0.0)
        gy = max(max(Y[iy0] - y, y - Y[iy1]), 0.0)
        lb = np.sqrt(gx * gx + gy * gy)
        ddx = X[ix0:ix1 + 1][None, :] - x
        ddy = Y[iy0:iy1 + 1][:, None] - y
        assert (np.sqrt(ddx * ddx + ddy * ddy) >= lb).all()
