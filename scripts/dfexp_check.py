"""Float32 emulation of exp_df / kf_val / df_mulf_round (gpismap_b200/csrc/common.cuh) against the reference's
double-precision formulas (cpp/src/covFnc.cpp:29-33): counts results that round to a different float."""
import math
import numpy as np
f32 = np.float32
def fma(a, b, c): return f32(np.float64(a) * np.float64(b) + np.float64(c))
def two_sum(a, b):
    s = f32(a + b); bb = f32(s - a); return s, f32(f32(a - f32(s - bb)) + f32(b - bb))
def fast_two_sum(a, b):
    s = f32(a + b); return s, f32(b - f32(s - a))
def df_mul(xh, xl, yh, yl):
    p = f32(xh * yh); e = fma(xh, yh, -p); e = f32(e + f32(f32(xh * yl) + f32(xl * yh))); return fast_two_sum(p, e)
def df_add(xh, xl, yh, yl):
    s, e = two_sum(xh, yh); e = f32(e + f32(xl + yl)); return fast_two_sum(s, e)
LN2_HI, LN2_MID, LN2_LO = f32(0.693145751953125), f32(1.42860677e-06), f32(5.49560397e-14)
C = [(f32(4.16666679e-02), f32(-1.24176347e-09)), (f32(1.66666672e-01), f32(-4.96705388e-09)), (f32(0.5), f32(0)), (f32(1), f32(0)), (f32(1), f32(0))]
T = [f32(1.38888889e-03), f32(1.98412698e-04), f32(2.48015873e-05), f32(2.75573192e-06)]
def exp_df(x):
    x = f32(x); k = f32(np.rint(f32(x * f32(1.44269504))))
    r1 = fma(-k, LN2_HI, x); p = f32(k * LN2_MID); pe = fma(k, LN2_MID, -p)
    rh, rl = two_sum(r1, -p); rl = f32(rl - f32(pe + f32(k * LN2_LO))); rh, rl = fast_two_sum(rh, rl)
    t = fma(fma(fma(T[3], rh, T[2]), rh, T[1]), rh, T[0])
    uh, ul = two_sum(f32(8.33333377e-03), f32(rh * t)); ul = f32(ul + f32(-4.34617203e-10))
    for ch, cl in C:
        uh, ul = df_mul(uh, ul, rh, rl); uh, ul = df_add(uh, ul, ch, cl)
    sc = f32(2.0 ** int(k)); return f32(uh * sc), f32(ul * sc)
def mulf_round(c, eh, el):
    p = f32(eh * c); e = fma(eh, c, -p); e = fma(el, c, e); return f32(p + e)
def kf_round(ar, eh, el):
    sh, sl = two_sum(f32(1), ar); ph, pl = df_mul(sh, sl, eh, el); return f32(ph + pl)
if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for scale, rmax, n in ((0.04, 0.25, 100000), (1.2, 12.0, 50000)):
        a = f32(math.sqrt(3.0) / scale); bad = 0; worst = 0.0
        for r in rng.uniform(1e-4, rmax, n).astype(np.float32):
            ar = f32(a * r); x = f32(-ar); eh, el = exp_df(x); ed = math.exp(float(x))
            worst = max(worst, abs(float(eh) + float(el) - ed) / ed)
            c = f32(a * a * f32(0.0123))
            bad += (f32((1.0 + float(ar)) * ed) != kf_round(ar, eh, el)) + (f32(float(c) * ed) != mulf_round(c, eh, el))
        print(f"scale {scale}: worst rel err {worst:.2e}, {bad} of {2*n} products round differently")
