"""ctypes wrapper over oracle/_ref/liboracle_port.so (hmmer_oracle.c, the scalar C restatement).

TEST INFRASTRUCTURE ONLY.  Takes a pyhmmer_b200.plan7.OptimizedProfile only as a *container* of the
node-major tables (they are validated against the reference separately); every score is computed by
the plain C code of hmmer_oracle.c.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "liboracle_port.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError("oracle/_ref/liboracle_port.so missing: run `make -C oracle port`")
        L = ctypes.CDLL(LIB)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        for f in ("oracle_ssv", "oracle_msv", "oracle_msv_full"):
            getattr(L, f).restype = ci
            getattr(L, f).argtypes = [vp, ci, ci, vp, ci, ci, ci, ci, ci, cf, ctypes.POINTER(cf)]
        L.oracle_vit.restype = ci
        L.oracle_vit.argtypes = [vp, ci, ci, vp, vp, ci, ci, ci, ci, cf, ctypes.POINTER(cf)]
        L.oracle_fwd.restype = ci
        L.oracle_fwd.argtypes = [vp, ci, ci, vp, vp, cf, cf, cf, ctypes.POINTER(cf)]
        L.oracle_ssv_longtarget.restype = ci
        L.oracle_ssv_longtarget.argtypes = [vp, ci, ci, vp, ci, ci, ci, ci, ci, cf, ci, ci, vp, vp]
        L.oracle_vit_longtarget.restype = ci
        L.oracle_vit_longtarget.argtypes = [vp, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
        L.oracle_null1.restype = cf
        L.oracle_null1.argtypes = [ci]
        L.oracle_bias.restype = cf
        L.oracle_bias.argtypes = [vp, ci, ci, vp]
        _lib = L
    return _lib


def _len_params(L):
    """L-dependent scalars, computed with the same libm the reference uses (via ctypes libm)."""
    libm = ctypes.CDLL("libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    libm.roundf.restype = ctypes.c_float
    libm.roundf.argtypes = [ctypes.c_float]
    f32 = np.float32
    scale_b = f32(3.0 / np.log(2.0))
    scale_w = f32(500.0 / np.log(2.0))
    v = -libm.roundf(float(scale_b * f32(libm.logf(float(f32(3.0) / f32(L + 3))))))
    tjb = 255 if v > 255.0 else int(v)
    pmove = f32(3.0) / (f32(L) + f32(3.0))
    w = libm.roundf(float(scale_w * f32(libm.logf(float(pmove)))))
    xw_move = 32767 if w >= 32767.0 else (-32768 if w <= -32768.0 else int(w))
    return tjb, xw_move, float(pmove)


class Port:
    def __init__(self, om):
        self.om = om
        d = om._desc
        self.M = om.M
        self.s = dict(tbm=d.tbm_b, tec=d.tec_b, base=d.base_b, bias=d.bias_b, scale_b=d.scale_b,
                      xwEm=d.xw[0][0], xwEl=d.xw[0][1], base_w=d.base_w, scale_w=d.scale_w, xfEm=d.xf[0][0], xfEl=d.xf[0][1])

    def _run(self, fn, codes, *args):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        sc = ctypes.c_float()
        st = fn(codes.ctypes.data, codes.size, self.M, *args, ctypes.byref(sc))
        return sc.value, st

    def ssv(self, codes):
        tjb, _, _ = _len_params(len(codes)); s = self.s
        return self._run(lib().oracle_ssv, codes, self.om.msv_cost.ctypes.data, s["tbm"], s["tec"], tjb, s["base"], s["bias"], s["scale_b"])

    def msv(self, codes):
        tjb, _, _ = _len_params(len(codes)); s = self.s
        return self._run(lib().oracle_msv, codes, self.om.msv_cost.ctypes.data, s["tbm"], s["tec"], tjb, s["base"], s["bias"], s["scale_b"])

    def vit(self, codes):
        _, xw_move, _ = _len_params(len(codes)); s = self.s
        return self._run(lib().oracle_vit, codes, self.om.vit_rsc.ctypes.data, self.om.vit_tsc.ctypes.data,
                         s["xwEm"], s["xwEl"], xw_move, s["base_w"], s["scale_w"])

    def fwd(self, codes):
        _, _, pmove = _len_params(len(codes)); s = self.s
        return self._run(lib().oracle_fwd, codes, self.om.fwd_rsc.ctypes.data, self.om.fwd_tsc.ctypes.data, s["xfEm"], s["xfEl"], pmove)

    def ssv_longtarget(self, codes, F1=0.02, cap=100000):
        """oracle_ssv_longtarget on one chunk: (windows [n,3] = start, model end, length; scores).  The threshold is the
        reference's formula (msvfilter.c:289-327) with the length model of the profile's max_length."""
        import math
        d = self.om._desc
        maxL = int(d.max_length)
        tjb, _, _ = _len_params(maxL)
        f32 = np.float32
        p1 = f32(maxL) / (f32(maxL) + f32(1.0))
        nullsc = f32(float(maxL) * math.log(float(p1)) + math.log(1.0 - float(p1)))            # p7_bg_NullOne (p7_bg.c:357)
        mu, lam = float(d.evparam[0]), float(d.evparam[1])
        invP = f32(mu - math.log(-1.0 * math.log(1.0 - F1)) / lam)                             # esl_gumbel_invsurv
        s = self.s
        thr = int(math.ceil((float(nullsc) + float(invP) * 0.69314718055994529 + 3.0) * float(f32(s["scale_b"])) + s["base"] + s["tec"] + tjb)) & 0xff
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        win = np.zeros((cap, 3), np.int64); wsc = np.zeros(cap, np.float32)
        n = lib().oracle_ssv_longtarget(codes.ctypes.data, codes.size, self.M, self.om.msv_cost.ctypes.data, s["tbm"], s["tec"], tjb,
                                        s["base"], s["bias"], s["scale_b"], thr, cap, win.ctypes.data, wsc.ctypes.data)
        assert n <= cap
        return win[:n].copy(), wsc[:n].copy()

    def vit_longtarget_threshold(self, cfg_len, filtersc, F2=3e-3):
        """The int16 score threshold p7_ViterbiFilter_longtarget derives from a P-value (vitfilter.c:330-346): float invP,
        double arithmetic inside ceil(), truncation to int16."""
        import math
        d = self.om._desc
        _, xw_move, _ = _len_params(cfg_len)
        mu, lam = float(d.evparam[2]), float(d.evparam[3])                                     # p7_VMU, p7_VLAMBDA
        log_part = (math.pow(F2, F2) - 1.0) / F2 if F2 < 5e-9 else math.log(-1.0 * math.log(1.0 - F2))   # esl_gumbel_invsurv
        invP = float(np.float32(mu - log_part / lam))
        s = self.s
        inner = (float(np.float32(filtersc)) + 0.69314718055994529 * invP + 3.0) * float(np.float32(s["scale_w"]))
        t = int(math.ceil(inner - float(s["xwEm"]) - float(xw_move) + float(s["base_w"])))
        return ((t + 32768) & 0xffff) - 32768, xw_move

    def vit_longtarget(self, codes, cfg_len, filtersc, F2=3e-3, cap=100000):
        """oracle_vit_longtarget on one window with the length model of <cfg_len>: landmarks [n,2] = i, k."""
        thr, xw_move = self.vit_longtarget_threshold(cfg_len, filtersc, F2)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        s = self.s
        hit = np.zeros((cap, 2), np.int64)
        n = lib().oracle_vit_longtarget(codes.ctypes.data, codes.size, self.M, self.om.vit_rsc.ctypes.data, self.om.vit_tsc.ctypes.data,
                                        s["xwEm"], s["xwEl"], xw_move, s["base_w"], int(self.om._desc.ddbound_w), thr, cap, hit.ctypes.data)
        assert n <= cap
        return hit[:n].copy()

    @staticmethod
    def null1(L):
        return lib().oracle_null1(int(L))
