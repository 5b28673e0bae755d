"""Test-only CPU interpreter of vp_op programs (numpy, fp64 accumulation).

It executes the exact op list + static memory plan + packed weight arena that the host lowers for the GPU, following
the op semantics documented in include/vpb200.h / DESIGN.md.  Used by the CPU test-suite to verify the lowering
(weight folding, permutations, buffer reuse, op parameters) against the oracle WITHOUT a GPU; the GPU tests then only
have the kernels left to prove.  Never imported by the product."""
import numpy as np

from mvector import _lib as L


def _act(v, a):
    if a == L.ACT_RELU:
        return np.maximum(v, 0)
    if a == L.ACT_HARDTANH20:
        return np.clip(v, 0, 20)
    if a == L.ACT_SIGMOID:
        return 1 / (1 + np.exp(-v))
    if a == L.ACT_TANH:
        return np.tanh(v)
    if a == L.ACT_SILU:
        return v / (1 + np.exp(-v))
    return v


class Sim:
    def __init__(self, pb, blob, feats, out_floats=None):
        self.ws = np.full(max(pb.peak, 256) // 4, np.nan, dtype=np.float32)   # NaN-poisoned: reads of stale data show
        self.blob = np.asarray(blob, dtype=np.float32)
        self.inp = np.asarray(feats, dtype=np.float32).reshape(-1)
        assert self.inp.size == pb.in_floats
        self.out = np.full(pb.out_floats, np.nan, dtype=np.float32)
        self.pb = pb

    def _mem(self, off):
        if off == L.BUF_INPUT:
            return self.inp, 0
        if off == L.BUF_OUTPUT:
            return self.out, 0
        assert off >= 0 and off % 16 == 0
        return self.ws, off // 4

    def rd(self, off, rows, ld, coff, C):
        mem, base = self._mem(off)
        idx = base + (np.arange(rows)[:, None] * ld + coff + np.arange(C)[None, :])
        return mem[idx].astype(np.float64)

    def wr(self, off, rows, ld, coff, C, val):
        mem, base = self._mem(off)
        idx = base + (np.arange(rows)[:, None] * ld + coff + np.arange(C)[None, :])
        mem[idx] = val.astype(np.float32)

    def w(self, off, n):
        assert off >= 0 and off % 16 == 0
        return self.blob[off // 4: off // 4 + n].astype(np.float64)

    def wmat(self, off, rows, ld, cols):
        idx = off // 4 + np.arange(rows)[:, None] * ld + np.arange(cols)[None, :]
        return self.blob[idx].astype(np.float64)

    def run(self):
        for o in self.pb.ops:
            {L.OP_CONV: self.conv, L.OP_CONV_C1: self.conv, L.OP_COLSTATS: self.colstats,
             L.OP_ASP_POOL: self.asp, L.OP_EW: self.ew, L.OP_POOL2D: self.pool2d}[o.kind](o)
        assert not np.isnan(self.out).any(), 'program output not fully written'
        return self.out.copy()

    # ------------------------------------------------------------------
    def conv(self, o):
        B = o.B
        cin_tot = o.Cin + (o.Cin2 if o.src2_mode == L.SRC2_CONCAT else 0)
        rows_in = B * o.Tin * o.Fin
        M = B * o.Tout * o.Fout
        X = self.rd(o.src, rows_in, o.in_ld, o.in_coff, o.Cin)
        if o.src2_mode == L.SRC2_ADD:
            X = X + self.rd(o.src2, rows_in, o.src2_ld, o.src2_coff, o.Cin)
        elif o.src2_mode == L.SRC2_CONCAT:
            X = np.concatenate([X, self.rd(o.src2, rows_in, o.src2_ld, o.src2_coff, o.Cin2)], axis=1)
        if o.pre_s >= 0:
            X = X * self.w(o.pre_s, cin_tot)[None] + self.w(o.pre_h, cin_tot)[None]
            if o.pre_relu:
                X = np.maximum(X, 0)
        K = o.KT * o.KF * cin_tot
        W = self.wmat(o.w, o.Cout, o.w_ld, K)
        m = np.arange(M)
        b = m // (o.Tout * o.Fout)
        to = (m // o.Fout) % o.Tout
        fo = m % o.Fout
        acc = np.zeros((M, o.Cout))
        for kt in range(o.KT):
            for kf in range(o.KF):
                ti = to * o.sT - o.padT + kt * o.dT
                fi = fo * o.sF - o.padF + kf * o.dF
                if o.pad_mode == L.PAD_REFLECT:
                    ti = np.where(ti < 0, -ti, ti)
                    ti = np.where(ti >= o.Tin, 2 * (o.Tin - 1) - ti, ti)
                ok = (ti >= 0) & (ti < o.Tin) & (fi >= 0) & (fi < o.Fin)
                row = (b * o.Tin + np.clip(ti, 0, o.Tin - 1)) * o.Fin + np.clip(fi, 0, o.Fin - 1)
                A = X[row] * ok[:, None]
                tap = kt * o.KF + kf
                acc += A @ W[:, tap * cin_tot:(tap + 1) * cin_tot].T
        seg = np.minimum(to // o.seg_len, o.n_seg - 1)
        urow = b * o.n_seg + seg
        v = acc
        if o.bias >= 0:
            v = v + self.w(o.bias, o.Cout)[None]
        if o.ubias != L.BUF_NONE:
            v = v + self.rd(o.ubias, B * o.n_seg, o.Cout, 0, o.Cout)[urow]
        v = _act(v, o.act)
        if o.post_s >= 0:
            v = v * self.w(o.post_s, o.Cout)[None] + self.w(o.post_h, o.Cout)[None]
        if o.gate != L.BUF_NONE:
            v = v * self.rd(o.gate, B * o.n_seg, o.Cout, 0, o.Cout)[urow]
        if o.res != L.BUF_NONE:
            v = v + self.rd(o.res, M, o.res_ld, o.res_coff, o.Cout)
        v = _act(v, o.act2)
        self.wr(o.dst, M, o.out_ld, o.out_coff, o.Cout, v)
        if o.sum != L.BUF_NONE:
            self.wr(o.sum, M, o.sum_ld, o.sum_coff, o.Cout, self.rd(o.sum, M, o.sum_ld, o.sum_coff, o.Cout) + v)

    def colstats(self, o):
        R = o.Tin * o.Fin
        X = self.rd(o.src, o.B * R, o.in_ld, o.in_coff, o.Cin).reshape(o.B, R, o.Cin)
        mean = X.mean(1)
        if o.mode == L.STATS_MEAN:
            self.wr(o.dst, o.B, o.out_ld, o.out_coff, o.Cin, mean)
            return
        if o.mode == L.STATS_SEG_CONTEXT:
            out = np.zeros((o.B, o.n_seg, o.Cin))
            for s in range(o.n_seg):
                out[:, s] = mean + X[:, s * o.seg_len:(s + 1) * o.seg_len].mean(1)
            self.wr(o.dst, o.B * o.n_seg, o.out_ld, o.out_coff, o.Cin, out.reshape(-1, o.Cin))
            return
        ssq = ((X - mean[:, None]) ** 2).sum(1)
        if o.mode == L.STATS_MEAN_STD_CLAMP:
            sd = np.sqrt(np.maximum(ssq / R, o.eps))
        elif o.mode == L.STATS_MEAN_STD_UNBIASED:
            sd = np.sqrt(ssq / (R - 1))
        elif o.mode == L.STATS_MEAN_VAR_UNBIASED:
            sd = ssq / (R - 1)
        else:
            sd = np.sqrt(ssq / (R - 1) + o.eps)
        self.wr(o.dst, o.B, o.out_ld, o.out_coff, 2 * o.Cin, np.concatenate([mean, sd], 1))

    def asp(self, o):
        T, Cn = o.Tin, o.Cin
        X = self.rd(o.src, o.B * T, o.in_ld, o.in_coff, Cn).reshape(o.B, T, Cn)
        Lg = self.rd(o.src2, o.B * T, o.src2_ld, o.src2_coff, Cn).reshape(o.B, T, Cn)
        e = np.exp(Lg - Lg.max(1, keepdims=True))
        a = e / e.sum(1, keepdims=True)
        mean = (a * X).sum(1)
        if o.mode == 1:
            self.wr(o.dst, o.B, o.out_ld, o.out_coff, Cn, mean)
            return
        sd = np.sqrt(np.maximum((a * (X - mean[:, None]) ** 2).sum(1), o.eps))
        self.wr(o.dst, o.B, o.out_ld, o.out_coff, 2 * Cn, np.concatenate([mean, sd], 1))

    def pool2d(self, o):
        X = self.rd(o.src, o.B * o.Tin * o.Fin, o.in_ld, o.in_coff, o.Cin).reshape(o.B, o.Tin, o.Fin, o.Cin)
        out = np.zeros((o.B, o.Tout, o.Fout, o.Cin))
        for to in range(o.Tout):
            for fo in range(o.Fout):
                t0, f0 = to * o.sT - o.padT, fo * o.sF - o.padF
                win = X[:, max(t0, 0):min(t0 + o.KT, o.Tin), max(f0, 0):min(f0 + o.KF, o.Fin)]
                out[:, to, fo] = win.max(axis=(1, 2)) if o.mode == 0 else win.sum(axis=(1, 2)) / (o.KT * o.KF)
        self.wr(o.dst, o.B * o.Tout * o.Fout, o.out_ld, o.out_coff, o.Cin, out.reshape(-1, o.Cin))

    def ew(self, o):
        rpu = o.Tin * o.Fin
        rows = o.B * rpu
        v = self.rd(o.src, rows, o.in_ld, o.in_coff, o.Cin)
        if o.mode == L.EW_PAD_COPY:
            self.wr(o.dst, rows, o.out_ld, o.out_coff, o.Cout, np.pad(v, ((0, 0), (0, o.Cout - o.Cin))))
            return
        if o.mode == L.EW_GATE_RES:
            if o.gate != L.BUF_NONE:
                v = v * self.rd(o.gate, o.B, o.Cin, 0, o.Cin)[np.arange(rows) // rpu]
            if o.res != L.BUF_NONE:
                v = v + self.rd(o.res, rows, o.res_ld, o.res_coff, o.Cin)
            v = _act(v, o.act2)
        elif o.mode == L.EW_AFF:
            y = self.rd(o.src2, rows, o.src2_ld, o.src2_coff, o.Cin)
            a = 1 + np.tanh(self.rd(o.res, rows, o.res_ld, o.res_coff, o.Cin))
            v = v * a + y * (2 - a)
        self.wr(o.dst, rows, o.out_ld, o.out_coff, o.Cin, v)


def simulate(model, feats):
    """model: a loaded Backbone mirror; feats [B, T, F] -> embeddings [B, embd] via the lowered program on CPU."""
    feats = np.asarray(feats, dtype=np.float32)
    B, T, _ = feats.shape
    pb = model.lower(B, T)
    out = Sim(pb, model._blob, feats).run()
    return out.reshape(B, model.embd_dim), pb
