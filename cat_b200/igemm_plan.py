"""Host-side description of the implicit GEMMs: K-unit tables and geometry descriptors.

Every dense convolution of the CAT path (and its data / weight gradients) is expressed as
    Y[row, n] = sum_units gather(X)[row, unit] . W[n, unit]
where a *unit* is 8 consecutive channels of one filter tap (include/catb200.h, catb_gather_unit /
catb_weight_unit).  This module builds the unit tables for
  * nn.Conv2d forward (zero or reflect padding, stride 1/2)        -> conv_fprop_units
  * its input gradient (optionally phase-decomposed for stride 2)   -> conv_dgrad_units
  * nn.ConvTranspose2d forward / input gradient                     -> convT_fprop_units / convT_dgrad_units
and K-concatenations of several convolutions that share an input or output buffer (the six branches
of InvertedResidualChannels, reference models/modules/inception_modules.py:124-180,230-236).

`emulate_fprop` / `emulate_wgrad` restate the device gather in plain torch so that the tables can be
validated against F.conv2d on CPU (tests/test_plan_cpu.py); they are test helpers, not a fallback.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import torch

PAD_ZERO, PAD_REFLECT = 0, 1


def cpad(c: int) -> int:
    """Channel count padded to the 16-byte unit (8 bf16)."""
    return (c + 7) // 8 * 8


@dataclass
class Units:
    """Parallel lists: gather side (dr, ds, cu) and weight side (w_off, sn_w, sc_w, nvalid)."""
    g: List[Tuple[int, int, int]] = field(default_factory=list)
    w: List[Tuple[int, int, int, int]] = field(default_factory=list)

    def __len__(self):
        return len(self.g)

    def extend(self, other: 'Units'):
        self.g.extend(other.g)
        self.w.extend(other.w)
        return self

    def phase(self, a: int, b: int) -> 'Units':
        """Units that contribute to output rows/cols of parity (a, b) when sd == 2, sn == 1."""
        out = Units()
        for gu, wu in zip(self.g, self.w):
            if (a + gu[0]) % 2 == 0 and (b + gu[1]) % 2 == 0:
                out.g.append(gu)
                out.w.append(wu)
        return out


def conv_fprop_units(w_off, Cout, Cin, R, S, pad, cu0=0, pad_s=None) -> Units:
    """nn.Conv2d weight [Cout, Cin, R, S] at arena offset w_off; input channels start at unit cu0 of
    the gathered buffer.  GEMM rows = output channels.  Used for fprop and (same tables) wgrad.
    pad_s: horizontal padding when it differs from the vertical one (x-packed 7x1 convs)."""
    pad_s = pad if pad_s is None else pad_s
    u = Units()
    for r in range(R):
        for s in range(S):
            for cu in range(cpad(Cin) // 8):
                u.g.append((r - pad, s - pad_s, cu0 + cu))
                u.w.append((w_off + cu * 8 * R * S + r * S + s, Cin * R * S, R * S, max(0, min(8, Cin - cu * 8))))
    return u


def tap_split_units(w_off, Cin, R, S, cu0=0) -> Units:
    """Conv weight [1, Cin, R, S] read as the 1x1 "tap-product" GEMM P[pix, t] = sum_c X[pix, c] W[0, c, t]: GEMM rows =
    the R*S taps, K = input channels.  The same table drives the weight gradient with dP as the lattice tensor."""
    u = Units()
    for cu in range(cpad(Cin) // 8):
        u.g.append((0, 0, cu0 + cu))
        u.w.append((w_off + cu * 8 * R * S, 1, R * S, max(0, min(8, Cin - cu * 8))))
    return u


def tap_split_dgrad_units(w_off, Cin, R, S, cu0=0) -> Units:
    """Input gradient of the tap-split conv: dX[pix, c] = sum_t dP[pix, t] W[0, c, t]: GEMM rows = input channels, K = taps."""
    u = Units()
    for tu in range(cpad(R * S) // 8):
        u.g.append((0, 0, cu0 + tu))
        u.w.append((w_off + tu * 8, R * S, 1, max(0, min(8, R * S - tu * 8))))
    return u


def conv_embedded_units(w_off, Cout, Cin, k, Kmax, cu0=0) -> Units:
    """A k x k conv (padding (k-1)/2) expressed on the tap grid of a Kmax x Kmax conv (padding (Kmax-1)/2):
    same gather side as conv_fprop_units(., ., Cin, Kmax, Kmax, (Kmax-1)/2), weight side pointing at the
    k x k tensor for the central taps and empty (nvalid = 0) for the outer ring.  Exact under reflect or
    zero padding because the outer taps carry zero weight.  Used to N-concatenate the first-stage convs
    of a residual block (kernel sizes 1/3/5 reading the same input) into one GEMM."""
    assert (Kmax - k) % 2 == 0 and k <= Kmax
    off, pad = (Kmax - k) // 2, (Kmax - 1) // 2
    u = Units()
    for r in range(Kmax):
        for s in range(Kmax):
            rb, sb = r - off, s - off
            inside = 0 <= rb < k and 0 <= sb < k
            for cu in range(cpad(Cin) // 8):
                u.g.append((r - pad, s - pad, cu0 + cu))
                if inside:
                    u.w.append((w_off + cu * 8 * k * k + rb * k + sb, Cin * k * k, k * k, max(0, min(8, Cin - cu * 8))))
                else:
                    u.w.append((0, 0, 0, 0))
    return u


def conv_dgrad_units(w_off, Cout, Cin, R, S, q, cu0=0, q_s=None) -> Units:
    """Input gradient of the same convolution: gather dY with offset dr = q - r (q = pad for a direct
    zero-padded gradient, q = 0 when the result is the gradient w.r.t. the *padded* frame of a reflect
    padded conv, folded afterwards by catb_reflect_fold).  GEMM rows = input channels."""
    q_s = q if q_s is None else q_s
    u = Units()
    for r in range(R):
        for s in range(S):
            for nu in range(cpad(Cout) // 8):
                u.g.append((q - r, q_s - s, cu0 + nu))
                u.w.append((w_off + nu * 8 * Cin * R * S + r * S + s, R * S, Cin * R * S, max(0, min(8, Cout - nu * 8))))
    return u


def conv_dgrad_embedded_units(w_off, Cout, Cin, k, Kmax, cu0=0) -> Units:
    """Input gradient of a k x k conv (reflect padding (k-1)/2, result = gradient w.r.t. its padded frame) expressed on
    the frame of a Kmax x Kmax conv: same gather side as conv_dgrad_units(., Cout, ., Kmax, Kmax, 0); the small conv's
    frame sits (Kmax - k) / 2 pixels inside the large one, so its tap (r', s') is the large tap (r' + off, s' + off) and
    the outer ring carries no weight.  Used to N-concatenate the last-conv input gradients of a residual block (one GEMM
    producing d(mid) of all six branches in the frame of the largest kernel, folded once)."""
    assert (Kmax - k) % 2 == 0 and k <= Kmax
    off = (Kmax - k) // 2
    u = Units()
    for r in range(Kmax):
        for s in range(Kmax):
            rb, sb = r - off, s - off
            inside = 0 <= rb < k and 0 <= sb < k
            for nu in range(cpad(Cout) // 8):
                u.g.append((-r, -s, cu0 + nu))
                if inside:
                    u.w.append((w_off + nu * 8 * Cin * k * k + rb * k + sb, k * k, Cin * k * k, max(0, min(8, Cout - nu * 8))))
                else:
                    u.w.append((0, 0, 0, 0))
    return u


def convT_fprop_units(w_off, Cin, Cout, R, S, pad, cu0=0) -> Units:
    """nn.ConvTranspose2d weight [Cin, Cout, R, S]: Y[oh] = sum X[(oh + pad - r)/stride] W[c, n, r]
    (use with sn=1, sd=stride).  GEMM rows = output channels."""
    u = Units()
    for r in range(R):
        for s in range(S):
            for cu in range(cpad(Cin) // 8):
                u.g.append((pad - r, pad - s, cu0 + cu))
                u.w.append((w_off + cu * 8 * Cout * R * S + r * S + s, R * S, Cout * R * S, max(0, min(8, Cin - cu * 8))))
    return u


def convT_dgrad_units(w_off, Cin, Cout, R, S, pad, cu0=0) -> Units:
    """Input gradient of the transposed conv = a strided correlation over dY (sn=stride, sd=1).
    GEMM rows = input channels.  The same tables drive its weight gradient (lattice tensor = X)."""
    u = Units()
    for r in range(R):
        for s in range(S):
            for nu in range(cpad(Cout) // 8):
                u.g.append((r - pad, s - pad, cu0 + nu))
                u.w.append((w_off + nu * 8 * R * S + r * S + s, Cout * R * S, R * S, max(0, min(8, Cout - nu * 8))))
    return u


def choose_n_tile(n_rows: int) -> int:
    """Rows of the packed weight image per CTA tile: a multiple of 16, at most 256, tiles balanced."""
    np16 = (n_rows + 15) // 16 * 16
    n_tiles = (np16 + 255) // 256
    per = (np16 + n_tiles - 1) // n_tiles
    return (per + 15) // 16 * 16


@dataclass
class Geometry:
    """The integer fields of catb_igemm_desc that do not depend on the unit table."""
    N: int
    H: int
    W: int
    ldx: int
    x_coff: int
    OH: int
    OW: int
    ldy: int
    y_coff: int
    sn: int = 1
    sd: int = 1
    pad_mode: int = PAD_ZERO
    o_step: int = 1
    o_ph: int = 0
    o_pw: int = 0

    @property
    def OHs(self):
        return (self.OH - self.o_ph + self.o_step - 1) // self.o_step

    @property
    def OWs(self):
        return (self.OW - self.o_pw + self.o_step - 1) // self.o_step


# ------------------------------------------------------------------------------------------------
# torch restatement of the device gather (test helper)
# ------------------------------------------------------------------------------------------------
def _reflect(i, L):
    i = i.abs()
    return torch.where(i >= L, 2 * (L - 1) - i, i)


def _gather_unit(x, geo: Geometry, gu, rows):
    """x: [N,H,W,ldx] float tensor; rows: (n, oh, ow) index tensors -> [M, 8] gathered values."""
    n, oh, ow = rows
    h = oh * geo.sn + gu[0]
    w = ow * geo.sn + gu[1]
    ok = torch.ones_like(h, dtype=torch.bool)
    if geo.sd == 2:
        ok &= (h % 2 == 0) & (w % 2 == 0)
        h = torch.div(h, 2, rounding_mode='floor')
        w = torch.div(w, 2, rounding_mode='floor')
    if geo.pad_mode == PAD_REFLECT:
        h, w = _reflect(h, geo.H), _reflect(w, geo.W)
    ok &= (h >= 0) & (h < geo.H) & (w >= 0) & (w < geo.W)
    hc, wc = h.clamp(0, geo.H - 1), w.clamp(0, geo.W - 1)
    c0 = geo.x_coff + gu[2] * 8
    vals = x[n, hc, wc, c0:c0 + 8]
    return vals * ok.unsqueeze(1).to(vals.dtype)


def _lattice_rows(geo: Geometry):
    n = torch.arange(geo.N).view(-1, 1, 1)
    oh = (geo.o_ph + torch.arange(geo.OHs) * geo.o_step).view(1, -1, 1)
    ow = (geo.o_pw + torch.arange(geo.OWs) * geo.o_step).view(1, 1, -1)
    n, oh, ow = torch.broadcast_tensors(n, oh, ow)
    return n.reshape(-1), oh.reshape(-1), ow.reshape(-1)


def emulate_fprop(geo: Geometry, units: Units, n_rows, x, arena, y, bias=None, accumulate=False):
    """y[N,OH,OW,ldy] (float) slice [y_coff : y_coff+n_rows] (+)= gather(x) . W^T on the sub-lattice."""
    rows = _lattice_rows(geo)
    acc = torch.zeros(rows[0].numel(), n_rows, dtype=x.dtype)
    ridx = torch.arange(n_rows)
    for gu, wu in zip(units.g, units.w):
        if wu[3] == 0:
            continue
        xg = _gather_unit(x, geo, gu, rows)[:, :wu[3]]
        q = torch.arange(wu[3])
        wmat = arena[wu[0] + ridx.view(-1, 1) * wu[1] + q.view(1, -1) * wu[2]]
        acc += xg @ wmat.T.to(x.dtype)
    if bias is not None:
        acc += bias[:n_rows].view(1, -1)
    n, oh, ow = rows
    if accumulate:
        acc += y[n, oh, ow, geo.y_coff:geo.y_coff + n_rows]
    y[n, oh, ow, geo.y_coff:geo.y_coff + n_rows] = acc
    return y


def emulate_wgrad(geo: Geometry, units: Units, n_rows, x, y, grad_arena):
    """grad_arena[w(row c, unit, q)] += sum_rows y[row, c] * gather(x)[row, unit*8+q]."""
    rows = _lattice_rows(geo)
    n, oh, ow = rows
    ymat = y[n, oh, ow, geo.y_coff:geo.y_coff + n_rows]
    ridx = torch.arange(n_rows)
    for gu, wu in zip(units.g, units.w):
        if wu[3] == 0:
            continue
        xg = _gather_unit(x, geo, gu, rows)[:, :wu[3]]
        contrib = ymat.T @ xg  # [n_rows, nvalid]
        q = torch.arange(wu[3])
        idx = wu[0] + ridx.view(-1, 1) * wu[1] + q.view(1, -1) * wu[2]
        grad_arena.index_put_((idx.reshape(-1),), contrib.reshape(-1).to(grad_arena.dtype), accumulate=True)
    return grad_arena


# ------------------------------------------------------------------------------------------------
# halo plan (v2 kernel): every input pixel is staged ONCE per CTA, taps address shifted windows
# ------------------------------------------------------------------------------------------------
@dataclass
class HaloPlan:
    """Metadata of catb_igemm_halo_fprop for one (Geometry, Units) pair.

    The lattice of an image is cut into vertical strips of TW columns; inside a strip, position (i, j) is
    flattened in *pitch space* m = i*Wf + j with Wf = TW + Xmax (columns j >= TW are garbage positions whose
    results are dropped).  A CTA owns `m_sub`*128 consecutive positions of one (image, strip) and stages, per
    64-channel chunk and parity plane p, the frame rows [m0, m0 + Lh), where frame pixel (fy, fx) of plane p
    is input pixel (mul*(fy + y0[p]) + pa[p], mul*(fx + x0[p] + strip*TW) + pb[p]).
    GEMM step s = (chunk, tap): A rows = halo rows plane*Lh + dy*Wf + dx + (m - m0); B = packed tile s."""
    units: Units                              # re-ordered, chunk aligned: 8 units per step
    steps: List[Tuple[int, int, int, int]]    # (chunk index, plane, dy, dx)
    chunks: List[Tuple[int, int, int, int]]   # (cu0, n_units, first_step, n_steps)
    planes: List[Tuple[int, int, int, int]]   # (pa, pb, y0, x0)
    mul: int
    OWs: int
    Ymax: int
    Xmax: int
    TW: int = 0
    m_sub: int = 1

    def __post_init__(self):
        if self.TW == 0:
            self.TW = self.OWs

    @property
    def n_strips(self):
        return (self.OWs + self.TW - 1) // self.TW

    @property
    def Wf(self):
        return self.TW + self.Xmax

    @property
    def Lh(self):
        return 128 * self.m_sub + self.Ymax * self.Wf + self.Xmax

    def halo_bytes(self):
        return len(self.planes) * self.Lh * 128


def make_halo_plan(geo: Geometry, units: Units):
    """Returns a HaloPlan, or None when the gather is not a shifted-window pattern."""
    if geo.sd == 2:
        if not (geo.sn == 1 and geo.o_step == 2):
            return None   # un-decomposed fractional stride: parity differs per output pixel
        eff = [((geo.o_ph + g[0]), (geo.o_pw + g[1])) for g in units.g]
        if any(a % 2 or b % 2 for a, b in eff):
            return None
        taps = [(a // 2, b // 2, 0, 0) for a, b in eff]
        mul = 1
    elif geo.sn == 2:
        if geo.o_step != 1:
            return None
        taps = [((g[0] - g[0] % 2) // 2, (g[1] - g[1] % 2) // 2, g[0] % 2, g[1] % 2) for g in units.g]
        mul = 2
    else:
        if geo.o_step != 1:
            return None
        taps = [(g[0], g[1], 0, 0) for g in units.g]
        mul = 1
    par = sorted({(t[2], t[3]) for t in taps})
    planes, Ymax, Xmax = [], 0, 0
    for (pa, pb) in par:
        ys = [t[0] for t in taps if (t[2], t[3]) == (pa, pb)]
        xs = [t[1] for t in taps if (t[2], t[3]) == (pa, pb)]
        planes.append((pa, pb, min(ys), min(xs)))
        Ymax, Xmax = max(Ymax, max(ys) - min(ys)), max(Xmax, max(xs) - min(xs))
    # Channel chunks: runs of up to 8 consecutive channel units that share the same tap set (so a
    # K-concatenation of convs with different kernel sizes is cut at the slice boundaries and no step
    # multiplies a slice by taps it does not have), each chunk = one staged halo of <= 64 channels.
    by_cu = {}
    for (gu, wu, t) in zip(units.g, units.w, taps):
        slot = by_cu.setdefault(gu[2], {})
        assert t not in slot, 'duplicate (channel unit, tap) in a halo GEMM'
        slot[t] = (gu, wu)
    groups = []
    for cu in sorted(by_cu):
        tapset = frozenset(by_cu[cu])
        if groups and len(groups[-1][1]) < 8 and groups[-1][1][-1] + 1 == cu and groups[-1][0] == tapset:
            groups[-1][1].append(cu)
        else:
            groups.append((tapset, [cu]))
    new_units, steps, chunks = Units(), [], []
    for tapset, cus in groups:
        first = len(steps)
        for t in sorted(tapset):
            pi = par.index((t[2], t[3]))
            steps.append((len(chunks), pi, t[0] - planes[pi][2], t[1] - planes[pi][3]))
            for j in range(8):
                if j < len(cus):
                    gu, wu = by_cu[cus[j]][t]
                    new_units.g.append(gu)
                    new_units.w.append(wu)
                else:  # padding unit: zero weights (points at a real unit: stays in bounds if ever gathered)
                    ref = by_cu[cus[0]][t][0]
                    new_units.g.append((ref[0], ref[1], ref[2]))
                    new_units.w.append((0, 0, 0, 0))
        chunks.append((cus[0], len(cus), first, len(steps) - first))
    return HaloPlan(new_units, steps, chunks, planes, mul, geo.OWs, Ymax, Xmax)


def emulate_halo_fprop(geo: Geometry, plan: HaloPlan, n_rows, x, arena, y, bias=None):
    """Torch restatement of the v2 device algorithm (halo fill + shifted windows), test helper."""
    Lh, Wf, TW = plan.Lh, plan.Wf, plan.TW
    M = 128 * plan.m_sub
    npos = geo.OHs * Wf
    Hf = geo.OHs + plan.Ymax
    ridx = torch.arange(n_rows)
    for n in range(geo.N):
        for strip in range(plan.n_strips):
            for m0 in range(0, npos, M):
                acc = torch.zeros(M, n_rows, dtype=x.dtype)
                for (cu0, n_units, first, nsteps) in plan.chunks:
                    halo = torch.zeros(len(plan.planes), Lh, 64, dtype=x.dtype)
                    h = torch.arange(Lh)
                    fy, fx = torch.div(m0 + h, Wf, rounding_mode='floor'), (m0 + h) % Wf
                    for pi, (pa, pb, y0, x0) in enumerate(plan.planes):
                        iy = plan.mul * (fy + y0) + pa
                        ix = plan.mul * (fx + x0 + strip * TW) + pb
                        if geo.pad_mode == PAD_REFLECT:
                            ok = (iy > -geo.H) & (iy < 2 * geo.H - 1) & (ix > -geo.W) & (ix < 2 * geo.W - 1)
                            iyr, ixr = _reflect(iy, geo.H), _reflect(ix, geo.W)
                        else:
                            iyr, ixr = iy, ix
                            ok = (iyr >= 0) & (iyr < geo.H) & (ixr >= 0) & (ixr < geo.W)
                        ok &= fy < Hf
                        c0 = geo.x_coff + cu0 * 8
                        vals = x[n, iyr.clamp(0, geo.H - 1), ixr.clamp(0, geo.W - 1), c0:c0 + n_units * 8]
                        halo[pi, :, :n_units * 8] = vals * ok.unsqueeze(1).to(x.dtype)
                    for s in range(first, first + nsteps):
                        _, plane, dy, dx = plan.steps[s]
                        a_off = dy * Wf + dx
                        A = halo[plane, a_off:a_off + M]                    # [M, 64]
                        Bm = torch.zeros(n_rows, 64, dtype=x.dtype)
                        for j in range(8):
                            wu = plan.units.w[s * 8 + j]
                            if wu[3]:
                                q = torch.arange(wu[3])
                                Bm[:, j * 8:j * 8 + wu[3]] = arena[wu[0] + ridx.view(-1, 1) * wu[1] + q.view(1, -1) * wu[2]]
                        acc += A @ Bm.T
                if bias is not None:
                    acc += bias[:n_rows].view(1, -1)
                m = m0 + torch.arange(M)
                i, j = torch.div(m, Wf, rounding_mode='floor'), m % Wf
                jg = strip * TW + j
                ok = (i < geo.OHs) & (j < TW) & (jg < geo.OWs)
                oh = geo.o_ph + i[ok] * geo.o_step
                ow = geo.o_pw + jg[ok] * geo.o_step
                y[n, oh, ow, geo.y_coff:geo.y_coff + n_rows] = acc[ok]
    return y
