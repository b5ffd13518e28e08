"""Golden-vector cases shared by make_golden.py (generator) and the tests.

Each case: geometry + input distribution + seed.  Inputs are regenerated from
the seed with oracle.dcnv3_oracle.make_inputs (CPU generator) and ALSO stored
in the fixture, so a change of torch's RNG cannot silently move them.
Sizes are kept small: fixtures are committed.
"""

def _c(name, N, H, W, G, gc, k=(3, 3), s=(1, 1), pad=(1, 1), dil=(1, 1), scale=1.0,
       dist="ref", seed=3, dtype="float32"):
    return dict(name=name, N=N, H=H, W=W, G=G, gc=gc, kh=k[0], kw=k[1], sh=s[0], sw=s[1],
                ph=pad[0], pw=pad[1], dh=dil[0], dw=dil[1], offset_scale=scale,
                dist=dist, seed=seed, dtype=dtype)


CASES = [
    # reference test fixture: models/ops_dcnv3/test.py:19-30 (fwd checks :33-90)
    _c("testpy_fwd_f32", 2, 8, 8, 4, 16, scale=2.0),
    _c("testpy_fwd_f64", 2, 8, 8, 4, 16, scale=2.0, dtype="float64"),
    # reference backward sweep: test.py:93-216 with N=2, M=2, D in {1,16,30,32,64,71}
    # (D=1025, test.py:257, is checked on the GPU against the oracle; too big to commit)
    _c("testpy_bwd_D1", 2, 8, 8, 2, 1, scale=2.0, seed=11),
    _c("testpy_bwd_D16", 2, 8, 8, 2, 16, scale=2.0, seed=12),
    _c("testpy_bwd_D30", 2, 8, 8, 2, 30, scale=2.0, seed=13),
    _c("testpy_bwd_D32", 2, 8, 8, 2, 32, scale=2.0, seed=14),
    _c("testpy_bwd_D64", 2, 8, 8, 2, 64, scale=2.0, seed=15),
    _c("testpy_bwd_D71", 2, 8, 8, 2, 71, scale=2.0, seed=16),
    _c("testpy_bwd_D16_f64", 2, 8, 8, 2, 16, scale=2.0, seed=17, dtype="float64"),
    # what the reference never tests: stride, dilation, 5x5, non-square, unit scale
    _c("unit_cfg1_small", 1, 14, 14, 4, 16, dist="unit", seed=0),
    _c("stride2_nonsquare", 2, 11, 13, 2, 8, s=(2, 2), scale=1.5, dist="unit", seed=21),
    _c("k5_dil2", 1, 12, 10, 2, 4, k=(5, 5), pad=(4, 4), dil=(2, 2), scale=1.0, dist="unit", seed=22),
    _c("k3_pad0", 1, 9, 9, 3, 8, pad=(0, 0), scale=1.0, dist="unit", seed=23),
    _c("far_offsets", 1, 8, 8, 2, 8, scale=4.0, dist="ref", seed=24),   # most points leave the map
]

BY_NAME = {c["name"]: c for c in CASES}

GEO_KEYS = ("kh", "kw", "sh", "sw", "ph", "pw", "dh", "dw")


def op_args(c):
    """Positional tail of dcnv3_core_pytorch / DCNv3Function after (input, offset, mask)."""
    return (c["kh"], c["kw"], c["sh"], c["sw"], c["ph"], c["pw"], c["dh"], c["dw"],
            c["G"], c["gc"], c["offset_scale"])
