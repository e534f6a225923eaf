"""include/pt_math.h and include/pt_glsl.h measured against independent references (no GPU).

The oracle and the kernels share these two headers, so a wrong pt_sin or mod() would be bit-identical on both sides and
pass every CPU<->GPU comparison.  Here the functions are checked against float64 libm / numpy -- the error bounds the
header states -- and, where oracle/_ref exists, against glm's implementation of the GLSL built-ins inside the
reference's own shader (tests/test_ref_pin.py does the latter on whole functions)."""
import numpy as np
import pytest

from oracle import oracle, sdf_build

N = 1 << 20


def ulp_err(got, ref64):
    """|got - ref| in units of the float32 spacing at |ref| (ref in float64)."""
    r32 = np.abs(ref64).astype(np.float32)
    ulp = np.spacing(np.maximum(r32, np.float32(1e-37))).astype(np.float64)
    return np.abs(got.astype(np.float64) - ref64) / ulp


@pytest.mark.parametrize('fn,name,lo,hi,bound', [
    (0, 'sin', -1e4, 1e4, 2.0), (1, 'cos', -1e4, 1e4, 2.0), (0, 'sin', -8.0, 8.0, 2.0), (1, 'cos', -8.0, 8.0, 2.0),
    (2, 'acos', -1.0, 1.0, 3.0), (3, 'exp2', -120.0, 120.0, 2.0), (5, 'exp', -80.0, 80.0, 3.0)])
def test_ulp_error_against_float64_libm(fn, name, lo, hi, bound):
    rng = np.random.default_rng(fn + 11)
    x = rng.uniform(lo, hi, N).astype(np.float32)
    got = oracle.math_eval(fn, x)
    ref = {'sin': np.sin, 'cos': np.cos, 'acos': np.arccos, 'exp2': np.exp2, 'exp': np.exp}[name](x.astype(np.float64))
    e = ulp_err(got, ref)
    if name in ('sin', 'cos'):
        # near a zero of the function the result is tiny and the error is that of the reduced argument: bound it in
        # absolute terms there (1e-7 = half an ulp of 1), in ulps elsewhere
        small = np.abs(ref) < 1e-2
        assert np.abs(got[small].astype(np.float64) - ref[small]).max() < 1.5e-7
        e = e[~small]
    if name == 'exp':
        # pt_exp(x) = exp2(x * log2 e) by definition, so the rounding of the product is amplified by
        # |x| ln 2: the Vulkan / GLSL precision requirement for exp() has exactly this shape, 3 + 2 |x| ulp
        assert (e <= 3.0 + 2.0 * np.abs(x.astype(np.float64))).all()
        e = e[np.abs(x) <= 1.0]
    print('pt_%s on [%g, %g]: max %.2f ulp, mean %.3f ulp' % (name, lo, hi, e.max(), e.mean()))
    assert e.max() <= bound


@pytest.mark.parametrize('fn,name,bound', [(4, 'log2', 3.0), (6, 'log', 3.0)])
def test_log_ulp_error(fn, name, bound):
    rng = np.random.default_rng(fn)
    x = np.exp2(rng.uniform(-120, 120, N)).astype(np.float32)
    got = oracle.math_eval(fn, x)
    ref = (np.log2 if name == 'log2' else np.log)(x.astype(np.float64))
    away = np.abs(ref) > 0.05   # the header's bound: "<= 3 ulp away from 1"
    e = ulp_err(got[away], ref[away])
    print('pt_%s: max %.2f ulp away from 1, mean %.3f' % (name, e.max(), e.mean()))
    assert e.max() <= bound
    near = ~away
    assert np.abs(got[near].astype(np.float64) - ref[near]).max() < 2e-8


def test_pow_relative_error():
    """pt_pow(x, y) = exp2(y log2 x) on the ranges the shader uses it on: Planck's l^-5 with l in metres,
    T^5, the cube roots of the cubic solver, 2^(-8/(FPS p))."""
    rng = np.random.default_rng(5)
    cases = [(rng.uniform(3.6e-7, 8e-7, N), np.full(N, -5.0)), (rng.uniform(1000, 12000, N), np.full(N, 5.0)),
             (np.exp2(rng.uniform(-30, 30, N)), np.full(N, 0.3333333)), (np.full(N, 2.0), rng.uniform(-20, 0, N)),
             (rng.uniform(0.0, 1.0, N), np.full(N, 1.0 / 2.4))]
    for x, y in cases:
        x, y = x.astype(np.float32), y.astype(np.float32)
        got = oracle.math_eval(7, x, y).astype(np.float64)
        ref = np.power(x.astype(np.float64), y.astype(np.float64))
        ok = ref > 0
        rel = np.abs(got[ok] / ref[ok] - 1.0)
        assert rel.max() < 2e-5, rel.max()   # |y log2 x| up to ~110: the product's rounding is amplified by ln 2 * 110


def test_special_values():
    x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan], dtype=np.float32)
    s, c = oracle.math_eval(0, x), oracle.math_eval(1, x)
    assert s[0] == 0 and s[1] == 0 and c[0] == 1 and c[1] == 1
    assert np.isnan(s[2:]).all() and np.isnan(c[2:]).all()
    e = oracle.math_eval(3, np.array([-200.0, 0.0, 1.0, 200.0, np.nan], dtype=np.float32))
    assert e[0] == 0 and e[1] == 1 and e[2] == 2 and np.isinf(e[3]) and np.isnan(e[4])
    lg = oracle.math_eval(4, np.array([0.0, 1.0, 2.0, -1.0, np.inf], dtype=np.float32))
    assert np.isneginf(lg[0]) and lg[1] == 0 and lg[2] == 1 and np.isnan(lg[3]) and np.isposinf(lg[4])
    a = oracle.math_eval(2, np.array([1.0, -1.0, 0.0, 1.5], dtype=np.float32))
    assert a[0] == 0 and abs(a[1] - np.pi) < 1e-6 and abs(a[2] - np.pi / 2) < 1e-6 and np.isnan(a[3])


# ---- pt_glsl.h: the GLSL built-ins the snippets see, against their GLSL 4.50 section 8 definitions in numpy -------------------

def glsl_eval(expr, pts):
    """Compile `float sdf(in vec3 p) { return <expr>; }` against include/pt_glsl.h with g++ (the oracle's build of the
    snippet hook) and evaluate it at pts."""
    import ctypes as C
    so = sdf_build.build(['float sdf(in vec3 p) { return %s; }\nfloat sdfmaterial(in vec3 p) { return 0.0; }' % expr])
    L = C.CDLL(so)
    L.oracle_SDF.restype = C.c_float
    L.oracle_SDF.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_uint]
    z = np.zeros(6, dtype=np.float32)
    return np.array([L.oracle_SDF(z.ctypes.data_as(C.c_void_p), float(x), float(y), float(w), 1) for x, y, w in pts], dtype=np.float32)


def test_glsl_builtins_against_numpy():
    rng = np.random.default_rng(9)
    pts = rng.uniform(-5, 5, (400, 3)).astype(np.float32)
    pts[:6] = [[0, 1, 0], [-0.0, 2, 1], [3, 3, 3], [-3, 2, 0.5], [2.5, -2, 1], [1e-20, 1, 0]]
    x, y, z = (pts[:, i] for i in range(3))
    f = np.float32
    with np.errstate(all='ignore'):
        exact = {
            'mod(p.x, p.y)': x - y * np.floor(x / y),                       # x - y * floor(x / y)
            'sign(p.x)': np.sign(x),
            'step(p.x, p.y)': np.where(y < x, f(0), f(1)),
            'mix(p.x, p.y, p.z)': x * (f(1) - z) + y * z,
            'clamp(p.x, -1.0, 2.0)': np.minimum(np.maximum(x, f(-1)), f(2)),
            'min(p.x, p.y)': np.where(y < x, y, x),
            'max(p.x, p.y)': np.where(x < y, y, x),
            'abs(p.x) + floor(p.y) + ceil(p.z)': np.abs(x) + np.floor(y) + np.ceil(z),
            'fract(p.x)': x - np.floor(x),
            'dot(p, vec3(1.0, 2.0, 3.0))': (x * f(1) + y * f(2)) + z * f(3),   # left-to-right accumulation
            'length(p)': np.sqrt((x * x + y * y) + z * z),
            'length(max(abs(p) - vec3(1.0), 0.0))': np.sqrt((np.maximum(np.abs(x) - 1, 0) ** 2 + np.maximum(np.abs(y) - 1, 0) ** 2) + np.maximum(np.abs(z) - 1, 0) ** 2),
            'normalize(p).y': y / np.sqrt((x * x + y * y) + z * z),
            'fma(p.x, p.y, p.z)': (x.astype(np.float64) * y.astype(np.float64) + z.astype(np.float64)).astype(f),
            'mod(p * 2.0, 2.0).z - 1.0': (z * 2 - f(2) * np.floor(z * 2 / f(2))) - 1,
            'inversesqrt(abs(p.x) + 1.0)': f(1) / np.sqrt(np.abs(x) + 1),
            'smin(p.x, p.y)': None, 'minMaterial(p.x, p.y, 3.0, 4.0)': np.where(x < y, f(3), f(4)),
        }
    for expr, want in exact.items():
        got = glsl_eval(expr, pts)
        if want is None:
            k = f(0.02) * f(6.0)
            h = np.maximum(k - np.abs(x - y), f(0)) / k
            want = np.where(y < x, y, x) - (h * h * h * f(0.5)) * k * f(0.3333333)
        want = want.astype(np.float32)
        both_nan = np.isnan(got) & np.isnan(want)
        assert np.array_equal(got[~both_nan].view(np.uint32), want[~both_nan].view(np.uint32)), expr


def test_glsl_transcendentals_within_tolerance():
    rng = np.random.default_rng(10)
    pts = rng.uniform(0.1, 4, (300, 3)).astype(np.float32)
    x, y = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)
    for expr, want in {'sin(p.x) + cos(p.y)': np.sin(x) + np.cos(y), 'pow(p.x, p.y)': x ** y, 'exp(-p.x) + log(p.y)': np.exp(-x) + np.log(y),
                       'sqrt(p.x) * acos(p.y / 4.5)': np.sqrt(x) * np.arccos(y / 4.5), 'exp2(p.x) - log2(p.y)': np.exp2(x) - np.log2(y)}.items():
        got = glsl_eval(expr, pts).astype(np.float64)
        assert np.abs(got - want).max() <= 2e-6 * np.maximum(np.abs(want), 1.0).max(), expr
