"""Record layout of the reference's star_t / binary_t (reference include/cmc/cmc.h:409-601, :231-390).

The arrays are kept byte for byte; the fields this path touches are exposed as numpy views.
"""
import numpy as np

SF_INFINITY = 1.0e10   # include/cmc/cmc.h:58
ZERO = 1.0e-20         # include/cmc/cmc.h:60
STAR_DTYPE_SIZE = 424
BINARY_DTYPE_SIZE = 464

STAR_OFFSETS = dict(r=0, vr=8, vt=16, m=24, E=32, J=40, EI=48, Eint=56, rnew=64, vrnew=72, vtnew=80, rOld=88, X=96, Y=104,
                    r_peri=112, r_apo=120, phi=128, interacted=136, threebb_interacted=144, binind=152, id=160, rad=168, Uoldrold=176, Uoldrnew=184, vtold=192, vrold=200, se_k=224)
STAR_INT_FIELDS = ("binind", "id", "interacted", "threebb_interacted")
STAR_INT32_FIELDS = ("se_k",)
BINARY_OFFSETS = dict(id1=0, id2=8, rad1=16, rad2=24, m1=32, m2=40, Eint1=48, Eint2=56, a=64, e=72, inuse=80)


def star_field(buf, name):
    """View of one star_t field over a (n, 424) uint8 array."""
    off = STAR_OFFSETS[name]
    dt = np.int64 if name in STAR_INT_FIELDS else (np.int32 if name in STAR_INT32_FIELDS else np.float64)
    flat = buf.reshape(-1)
    return np.lib.stride_tricks.as_strided(flat[off:off + np.dtype(dt).itemsize].view(dt), shape=(buf.shape[0],), strides=(STAR_DTYPE_SIZE,))


def binary_field(buf, name, dtype=None):
    off = BINARY_OFFSETS[name]
    dt = dtype or (np.int64 if name in ("id1", "id2") else (np.int32 if name == "inuse" else np.float64))
    flat = buf.reshape(-1)
    sz = np.dtype(dt).itemsize
    return np.lib.stride_tricks.as_strided(flat[off:off + sz].view(dt), shape=(buf.shape[0],), strides=(BINARY_DTYPE_SIZE,))


# Row of the HDF5 snapshot table (reference include/cmc/cmc.h:136-200 `Snapshot`, field names as write_snapshot() gives them to
# H5TBmake_table, src/cmc/cmc_io.c:2962-2970): 62 fields, 488 bytes (4 bytes of padding behind startype).
SNAPSHOT_FIELDS = ("id", "m_MSUN", "r", "vr", "vt", "E", "J", "binflag", "m0_MSUN", "m1_MSUN", "id0", "id1", "a_AU", "e", "startype", "luminosity_LSUN",
                   "radius_RSUN", "bin_startype0", "bin_startype1", "bin_star_lum0_LSUN", "bin_star_lum1_LSUN", "bin_star_radius0_RSUN",
                   "bin_star_radius1_RSUN", "bin_Eb", "eta", "star_phi", "rad0", "rad1", "tb", "lum0", "lum1", "massc0", "massc1", "radc0", "radc1",
                   "menv0", "menv1", "renv0", "renv1", "tms0", "tms1", "dmdt0", "dmdt1", "radrol0", "radrol1", "ospin0", "ospin1", "B0", "B1",
                   "formation0", "formation1", "bacc0", "bacc1", "tacc0", "tacc1", "mass0_0", "mass0_1", "epoch0", "epoch1", "ospin", "B", "formation")
_SNAP_LONG = {"id", "binflag", "id0", "id1"}
_SNAP_INT = {"startype", "bin_startype0", "bin_startype1"}
SNAPSHOT_DTYPE = np.dtype([(n, np.int64 if n in _SNAP_LONG else (np.int32 if n in _SNAP_INT else np.float64)) for n in SNAPSHOT_FIELDS], align=True)
assert SNAPSHOT_DTYPE.itemsize == 488
