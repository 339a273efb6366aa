"""Synthetic workloads of SURVEY.md section 8(d): databases, receiver arrays and candidate lists.

Everything here is deterministic input generation (numpy); no engine arithmetic.
"""
import numpy as np

from .engine import Gfdb, KIWIBENCH_STF

EARTH_RADIUS = 6371000.0


def bench_s_db(nx=200, nz=200):
    """kiwibench database (benchmark/kiwibench.py:45-91): gfdb_build benchdb 1 200 200 10 0.1 50 50 50 0."""
    return Gfdb.create(nx, nz, 10, 0.1, 50.0, 50.0, 50.0, 0.0).build_ahfull(2300.0, 3200.0, 1600.0, KIWIBENCH_STF)


def bench_l_db(nx=2000, nz=150, dx=100.0, dz=200.0, dt=0.1):
    """bench-L: 0.1-200 km x 0-29.8 km, rho=2700 alpha=6000 beta=3464, HBM-resident (> L2)."""
    return Gfdb.create(nx, nz, 10, dt, dx, dz, dx, 0.0).build_ahfull(2700.0, 6000.0, 3464.0, KIWIBENCH_STF)


def destination(lat0_deg, lon0_deg, dist_m, azi_rad):
    """Inverse great-circle problem on the sphere (fp64)."""
    lat0, lon0 = np.radians(lat0_deg), np.radians(lon0_deg)
    d = np.asarray(dist_m, dtype=np.float64) / EARTH_RADIUS
    lat = np.arcsin(np.sin(lat0) * np.cos(d) + np.cos(lat0) * np.sin(d) * np.cos(azi_rad))
    lon = lon0 + np.arctan2(np.sin(azi_rad) * np.sin(d) * np.cos(lat0), np.cos(d) - np.sin(lat0) * np.sin(lat))
    return np.degrees(lat), np.degrees(lon)


def receivers(n, origin=(30.0, 70.0), dmin=45e3, dmax=150e3, seed=12345):
    """n receivers at uniform epicentral distance / azimuth around the origin."""
    rng = np.random.default_rng(seed)
    dist = rng.uniform(dmin, dmax, n)
    azi = rng.uniform(0.0, 2.0 * np.pi, n)
    lat, lon = destination(origin[0], origin[1], dist, azi)
    return lat, lon, np.zeros(n, dtype=np.float32)


# the documented Izmit example (minimizer.f90:1632): bilateral time north east depth moment strike dip
# rake rupture-direction length-a length-b width rupture-velocity rise-time
IZMIT = np.array([0, 0, 0, 10000, 2e20, 91, 87, 164, 0, 40000, 20000, 18000, 3500, 2], dtype=np.float32)


def bilateral_sweep(n, base=IZMIT, seed=None):
    """Candidate list of config C5: strike/dip/rake +-10 deg, depth 10-14 km, length-a 30-50 km.
    n = k**5 gives the full k^5 lattice; otherwise the first n lattice points of the next larger k."""
    k = max(1, int(np.ceil(n ** (1.0 / 5.0) - 1e-9)))
    ax = lambda lo, hi: np.linspace(lo, hi, k) if k > 1 else np.array([(lo + hi) / 2.0])
    strike = base[5] + ax(-10, 10); dip = base[6] + ax(-10, 10) * 0.3; rake = base[7] + ax(-10, 10)
    depth = ax(10000, 14000); la = ax(30000, 50000)
    dip = np.clip(dip, 1.0, 90.0)
    grid = np.stack(np.meshgrid(strike, dip, rake, depth, la, indexing="ij"), -1).reshape(-1, 5)[:n]
    p = np.tile(base, (grid.shape[0], 1)).astype(np.float32)
    p[:, 5] = grid[:, 0]; p[:, 6] = grid[:, 1]; p[:, 7] = grid[:, 2]; p[:, 3] = grid[:, 3]; p[:, 9] = grid[:, 4]
    return p


def fibonacci_moment_tensors(n):
    """n unit moment tensors (mxx myy mzz mxy mxz myz): double couples with fault normal / slip on a
    Fibonacci sphere."""
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + 5 ** 0.5) * i
    nrm = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], -1)
    ref = np.where(np.abs(nrm[:, 2:3]) < 0.9, np.array([[0, 0, 1.0]]), np.array([[1.0, 0, 0]]))
    slip = np.cross(nrm, ref); slip /= np.linalg.norm(slip, axis=1, keepdims=True)
    m = nrm[:, :, None] * slip[:, None, :] + slip[:, :, None] * nrm[:, None, :]
    return np.stack([m[:, 0, 0], m[:, 1, 1], m[:, 2, 2], m[:, 0, 1], m[:, 0, 2], m[:, 1, 2]], -1).astype(np.float32)


def moment_tensor_sweep(nloc_side=10, nmt=100, risetime=1.0, extent=5000.0, depth=(8000.0, 12000.0)):
    """Config C2: (north, east, depth) lattice x unit tensors -> [nloc^3 * nmt, 11] moment_tensor params
    (time north east depth mxx myy mzz mxy mxz myz rise-time)."""
    ax = np.linspace(-extent, extent, nloc_side) if nloc_side > 1 else np.array([0.0])
    dz = np.linspace(depth[0], depth[1], nloc_side) if nloc_side > 1 else np.array([sum(depth) / 2])
    loc = np.stack(np.meshgrid(ax, ax, dz, indexing="ij"), -1).reshape(-1, 3)
    mts = fibonacci_moment_tensors(nmt) * 1e18
    p = np.zeros((loc.shape[0], nmt, 11), dtype=np.float32)
    p[:, :, 1:4] = loc[:, None, :]
    p[:, :, 4:10] = mts[None, :, :]
    p[:, :, 10] = risetime
    return p.reshape(-1, 11)
