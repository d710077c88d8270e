"""Thin Python wrappers over the C ABI: torch tensors in, raw pointers + current stream out.

PyTorch is used only for device memory and streams; every computation below is a kernel of
libcatb200.so.  All wrappers enqueue on ``torch.cuda.current_stream()`` and never synchronise, so a
sequence of calls can be captured into a CUDA graph.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _C
from .igemm_plan import Geometry, Units, choose_n_tile, cpad, make_halo_plan

ACT = {'none': _C.ACT_NONE, 'relu': _C.ACT_RELU, 'leaky': _C.ACT_LEAKY02, 'tanh': _C.ACT_TANH, 'leaky001': _C.ACT_LEAKY001}
BF16 = torch.bfloat16


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def require_cuda():
    if not torch.cuda.is_available():
        raise _C.CatbError('cat_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    _C.init(torch.cuda.current_device())


class Act:
    """A view of an NHWC bf16 activation buffer: channel slice [coff, coff+C) of pitch ld."""
    __slots__ = ('t', 'N', 'H', 'W', 'ld', 'coff', 'C')

    def __init__(self, t, coff=0, C=None):
        assert t.dtype == BF16 and t.dim() == 4 and t.is_contiguous()
        self.t = t
        self.N, self.H, self.W, self.ld = t.shape
        self.coff = coff
        self.C = self.ld - coff if C is None else C
        assert self.coff % 8 == 0 and self.C % 8 == 0 and self.coff + self.C <= self.ld

    @staticmethod
    def empty(N, H, W, C, device, zero=False):
        Cp = cpad(C)
        t = (torch.zeros if zero else torch.empty)(N, H, W, Cp, dtype=BF16, device=device)
        return Act(t)

    def slice(self, coff, C):
        return Act(self.t, self.coff + coff, C)

    @property
    def pixels(self):
        return self.N * self.H * self.W

    @property
    def HW(self):
        return self.H * self.W

    def args(self):
        return _p(self.t), self.ld, self.coff


_G_DT = np.dtype([('dr', 'i1'), ('ds', 'i1'), ('cu', '<i2')])


def units_to_device(units: Units, device):
    g = np.array(units.g, dtype=np.int64).reshape(-1, 3)
    ga = np.zeros(len(units), dtype=_G_DT)
    assert np.abs(g[:, :2]).max(initial=0) < 127 and g[:, 2].max(initial=0) < 32767
    ga['dr'], ga['ds'], ga['cu'] = g[:, 0], g[:, 1], g[:, 2]
    wa = np.array(units.w, dtype=np.int32).reshape(-1, 4)
    gt = torch.from_numpy(ga.view(np.int32).copy()).to(device)
    wt = torch.from_numpy(wa.copy()).to(device)
    return gt, wt


USE_HALO = os.environ.get('CATB_NO_HALO', '0') != '1'   # v2 (halo) forward kernel unless disabled
AUTOTUNE = os.environ.get('CATB_NO_AUTOTUNE', '0') != '1'  # pick v1 / v2 per GEMM by timing the first call
USE_PERSIST = os.environ.get('CATB_NO_PERSIST', '0') != '1'   # v3 (persistent halo kernel) variants offered to the autotune
FUSE_STATS = os.environ.get('CATB_NO_FUSED_STATS', '0') != '1'   # norm statistics accumulated by the conv epilogue
USE_TMA = os.environ.get('CATB_NO_TMA', '0') != '1'
TMA_REFLECT = os.environ.get('CATB_TMA_REFLECT', '0') == '1'   # offer v3 mode 3 (TMA + reflection fringe pass) to the autotune
TAP_HEAD = os.environ.get('CATB_NO_TAP_HEAD', '0') != '1'     # one-output-channel convs (PatchGAN head) in tap-split form           # v3 stages zero-padded halos with cp.async.bulk.tensor


_SCRATCH = {}


def _scratch_like(t):
    """A reusable zero-initialised scratch tensor shaped like `t` (autotuning of accumulating kernels)."""
    key = (t.device, t.numel(), t.dtype)
    if key not in _SCRATCH:
        _SCRATCH.clear()
        _SCRATCH[key] = torch.zeros_like(t)
    return _SCRATCH[key]


_UNPACK = [None]      # the UnpackQueue of the backward pass that is being issued (None: second stages run immediately)


class UnpackQueue:
    """Second stage of every two-stage weight gradient of one backward pass as ONE launch.

    ``with net.unpack_queue:`` around a backward pass makes Gemm.wgrad launch only its first stage (partial tiles into
    the GEMM's own workspace) and register the second stage here; leaving the block -- after every side stream has been
    joined -- runs them all through catb_wgrad_unpack_batch.  The job table lives on the device and is rebuilt only when
    the set of jobs changes (never during graph capture: the eager tuning step of an engine builds it first)."""

    def __init__(self, device):
        self.dev, self.jobs, self.tables, self.depth, self._outer = device, [], {}, 0, None

    def __enter__(self):
        if self.depth == 0:
            self._outer, _UNPACK[0] = _UNPACK[0], self
            self.jobs = []
        self.depth += 1
        return self

    def __exit__(self, exc_type, exc, tb):
        self.depth -= 1
        if self.depth == 0:
            _UNPACK[0] = self._outer
            if exc_type is None:
                self.flush()
        return False

    def add(self, ws, splits, ws_rows, ws_k, row0, n_rows, n_units, wt, grad):
        self.jobs.append((ws.data_ptr(), wt.data_ptr(), grad.data_ptr(), splits, ws_rows, ws_k, row0, n_rows, n_units))

    def flush(self):
        """Run (and forget) the jobs registered so far; called implicitly when the block is left."""
        if not self.jobs:
            return
        if str(self.dev) == 'cpu':       # kernel emulation: second stages were applied immediately
            self.jobs = []
            return
        key = tuple(self.jobs)
        entry = self.tables.get(key)
        if entry is None:
            assert not torch.cuda.is_current_stream_capturing(), 'unpack table must be built before graph capture'
            arr = (_C.UnpackJob * len(key))()
            big = 1
            for i, j in enumerate(key):
                arr[i] = _C.UnpackJob(*j)
                big = max(big, j[7] * j[8] * 8)
            table = torch.from_numpy(np.frombuffer(arr, dtype=np.uint8).copy()).to(self.dev)
            entry = (table, len(key), max(1, min(148 * 4, (big + 2047) // 2048)))
            self.tables[key] = entry
        table, n, blocks = entry
        _C.call('catb_wgrad_unpack_batch', _p(table), n, blocks, _stream())
        self.jobs = []


def flush_unpack():
    """Run the second stages registered so far (a consumer of the gradients follows inside the same backward pass)."""
    if _UNPACK[0] is not None:
        _UNPACK[0].flush()


class Gemm:
    """One implicit GEMM: geometry + unit tables (+ packed bf16 weights for the fprop direction).

    The forward direction runs the v2 halo kernel whenever the gather is a shifted-window pattern and the
    halo tile fits in shared memory (make_halo_plan / catb_igemm_halo_fits), else the v1 gather-per-tap
    kernel.  Both read the same packed weights; wgrad always uses the original unit order."""

    def __init__(self, geo: Geometry, units: Units, n_rows: int, device, need_pack=True, halo=None, force_tile=None,
                 segments=None, force_mode=None):
        """force_tile=(TW, m_sub) pins the halo tiling (tests); by default the widest strip / largest
        sub-tile count that fits in shared memory and still fills the GPU is chosen.
        segments=[(row0, span, nreal, Units)] describes an N-concatenation: every segment shares the gather
        side of `units` and supplies its own weight side for image rows [row0, row0+span).
        force_mode pins the forward kernel of the halo tilings (tests): 0 = v2 (one tile per CTA), 1 = v3 persistent with
        cp.async producers, 2 = v3 persistent with TMA-staged tiles, 3 = v3 with TMA tiles + the reflection fringe pass; by
        default every applicable one is a candidate."""
        assert len(units) > 0 and n_rows > 0
        self.segments = None
        self.seg_raw = segments      # also drives the per-segment second stage of the weight gradient
        self.geo, self.units, self.n_rows = geo, units, n_rows
        self.n_units = len(units)
        self.n_tile = choose_n_tile(n_rows)
        self.gt, self.wt = units_to_device(units, device)
        self.packed = None
        self.halo = None
        self.choice = None      # 'v1' | 'v2' once tuned; None = v2 whenever a halo plan exists
        self.tuned_ms = None
        self.w_ready, self.w_halo, self.w_choice, self.w_tuned_ms = False, None, None, None
        self.f_units, self.f_gt, self.f_wt = units, self.gt, self.wt   # tables of the forward direction
        if need_pack:
            lib = _C.load()
            plan = make_halo_plan(geo, units) if (USE_HALO if halo is None else halo) else None
            self.tilings = []   # [(TW, m_sub, HaloDesc, steps tensor, mode)]: candidates, the autotune keeps one
            if plan is not None:
                if force_tile is not None:
                    cands = [force_tile]
                else:   # widest strip that fits, with 2 and 1 sub-tiles; one narrower strip as alternative
                    widths = [geo.OWs] + [t for t in (64, 32, 16) if t < geo.OWs]
                    cands = [(tw, ms) for tw in widths for ms in ((4, 2, 1) if self.n_tile <= 128 else (2, 1))]
                n_strip_widths = 0
                # short weight tiles (thin GEMMs): a second variant with a small weight ring, so that 2-4 CTAs share
                # an SM and the fill / MMA / epilogue phases of neighbouring tiles overlap
                budgets = (0, 16 * 1024) if (force_tile is None and self.n_tile <= 128) else (0,)
                # v3 stages the halo with TMA where out-of-bounds = zero is the padding rule (or no tap leaves the image)
                no_border = plan.Ymax == 0 and plan.Xmax == 0 and all(pl[2] == 0 and pl[3] == 0 for pl in plan.planes)
                tma_ok = USE_TMA and (geo.pad_mode != _C.PAD_REFLECT or no_border)
                # reflection padding with border taps: TMA boxes + a fringe pass that mirrors the out-of-image halo rows (mode 3)
                tma_reflect = USE_TMA and not tma_ok
                self.c_visible = 8 * max(cu0 + nu for (cu0, nu, _, _) in plan.chunks)
                for tw, ms in cands:
                    plan.TW, plan.m_sub = tw, ms
                    modes = []
                    if lib.catb_igemm_halo_fits(len(plan.planes), plan.Lh, self.n_tile, ms, len(plan.steps), len(plan.chunks)):
                        modes.append(0)
                    if USE_PERSIST or force_mode:
                        if force_mode:
                            m3s = (force_mode,) if (force_mode == 1 or (force_mode == 2 and tma_ok) or
                                                    (force_mode == 3 and tma_reflect)) else ()
                        else:       # zero padding: TMA (else cp.async); reflection: cp.async producers -- the TMA + fringe-pass
                            # form lost to them and to v2 on every shape measured (profiles/r02_gemm_variants_v6.txt: the
                            # serial box -> fringe -> MMA chain adds latency to GEMMs that are MMA-issue bound anyway), so it
                            # is only a candidate when asked for (CATB_TMA_REFLECT=1)
                            m3s = (2, 1) if tma_ok else ((3, 1) if (tma_reflect and TMA_REFLECT) else (1,))
                        for m3 in m3s:
                            if lib.catb_igemm_halo_persist_fits(len(plan.planes), plan.Lh, plan.Wf, plan.mul, self.n_tile, ms,
                                                                len(plan.steps), len(plan.chunks), 0, int(m3 >= 2)):
                                modes.append(m3)
                                if m3 == 2:
                                    break
                    if force_mode is not None:
                        modes = [m for m in modes if m == force_mode]
                    if not modes:
                        continue
                    if force_tile is None and tw not in [t[0] for t in self.tilings]:
                        n_strip_widths += 1
                        if n_strip_widths > (3 if self.n_tile <= 128 else 2):
                            break
                    st = np.array([[pl * plan.Lh + dy * plan.Wf + dx, ci] for (ci, pl, dy, dx) in plan.steps], dtype=np.int32)
                    st = torch.from_numpy(st).to(device)
                    for mode in modes:
                        for budget in budgets:
                            hd = _C.HaloDesc()
                            hd.n_steps, hd.n_chunks, hd.n_planes = len(plan.steps), len(plan.chunks), len(plan.planes)
                            for i, (pa, pb, y0, x0) in enumerate(plan.planes):
                                hd.plane_pa[i], hd.plane_pb[i], hd.plane_y0[i], hd.plane_x0[i] = pa, pb, y0, x0
                            hd.mul, hd.TW, hd.n_strips, hd.Wf, hd.Lh = plan.mul, plan.TW, plan.n_strips, plan.Wf, plan.Lh
                            hd.Ymax, hd.Xmax, hd.m_sub, hd.b_budget = plan.Ymax, plan.Xmax, plan.m_sub, budget
                            self.tilings.append((tw, ms, hd, st, mode))
                if not self.tilings:
                    plan = None
            if plan is not None:
                self.halo = plan
                self.f_units = plan.units
                self.f_gt, self.f_wt = units_to_device(plan.units, device)
                self.h_chunks = torch.from_numpy(np.array(plan.chunks, dtype=np.int32)).to(device)
                # default before tuning: largest sub-tile count that still gives >= 2 waves of CTAs
                n_tiles_n = (n_rows + self.n_tile - 1) // self.n_tile
                pick = self.tilings[-1]
                for t in self.tilings:
                    wf = t[0] + plan.Xmax
                    ctas = geo.N * ((geo.OWs + t[0] - 1) // t[0]) * ((geo.OHs * wf + 128 * t[1] - 1) // (128 * t[1])) * n_tiles_n
                    if t[1] == 1 or ctas >= 2 * 148:
                        pick = t
                        break
                self._use_tiling(pick)
            # v1 reads the compact original table, v2 the chunk-aligned one: two packed images until tuned
            self.packed_v1 = torch.zeros(lib.catb_packed_weight_bytes(n_rows, self.n_units, self.n_tile),
                                         dtype=torch.uint8, device=device)
            if self.halo is not None:
                self.packed = torch.zeros(lib.catb_packed_weight_bytes(n_rows, len(self.f_units), self.n_tile),
                                          dtype=torch.uint8, device=device)
            if segments is not None:
                self.segments = []
                for (row0, span, nreal, su) in segments:
                    assert su.g == units.g, 'segments must share the gather side of the GEMM'
                    wt1 = units_to_device(su, device)[1]
                    wt2 = units_to_device(make_halo_plan(geo, su).units, device)[1] if self.halo is not None else None
                    self.segments.append((row0, span, nreal, wt1, wt2))

    def _use_tiling(self, t):
        tw, ms, hd, steps, mode = t
        self.halo.TW, self.halo.m_sub = tw, ms
        self.hdesc, self.h_steps, self.h_mode = hd, steps, mode

    def desc(self, act=0, accumulate=False, y_is_f32=False, geo=None, n_units=None):
        g = geo or self.geo
        d = _C.IgemmDesc()
        d.N, d.H, d.W, d.ldx, d.x_coff = g.N, g.H, g.W, g.ldx, g.x_coff
        d.OH, d.OW, d.ldy, d.y_coff = g.OH, g.OW, g.ldy, g.y_coff
        d.o_step, d.o_ph, d.o_pw, d.OHs, d.OWs = g.o_step, g.o_ph, g.o_pw, g.OHs, g.OWs
        d.sn, d.sd, d.pad_mode = g.sn, g.sd, g.pad_mode
        d.n_units, d.n_rows, d.n_tile = (self.n_units if n_units is None else n_units), self.n_rows, self.n_tile
        d.act, d.accumulate, d.y_is_f32 = int(act), int(bool(accumulate)), int(bool(y_is_f32))
        return d

    def pack_jobs(self):
        """The (wunits, packed image, n_tile, n_units, row0, span, nreal) tuples pack() would launch one by one."""
        do1 = self.halo is None or self.choice != 'v2'
        do2 = self.halo is not None and self.choice != 'v1'
        n_img = ((self.n_rows + self.n_tile - 1) // self.n_tile) * self.n_tile
        jobs = []
        if self.segments is None:
            if do1:
                jobs.append((self.wt, self.packed_v1, self.n_tile, self.n_units, 0, n_img, self.n_rows))
            if do2:
                jobs.append((self.f_wt, self.packed, self.n_tile, len(self.f_units), 0, n_img, self.n_rows))
            return jobs
        for (row0, span, nreal, wt1, wt2) in self.segments:
            if do1:
                jobs.append((wt1, self.packed_v1, self.n_tile, self.n_units, row0, span, nreal))
            if do2:
                jobs.append((wt2, self.packed, self.n_tile, len(self.f_units), row0, span, nreal))
        return jobs

    def pack(self, arena):
        do1 = self.halo is None or self.choice != 'v2'
        do2 = self.halo is not None and self.choice != 'v1'
        d1, d2 = self.desc(), self.desc(n_units=len(self.f_units))
        if self.segments is None:
            if do1:
                _C.call('catb_pack_weights', C.byref(d1), _p(self.wt), _p(arena), _p(self.packed_v1), _stream())
            if do2:
                _C.call('catb_pack_weights', C.byref(d2), _p(self.f_wt), _p(arena), _p(self.packed), _stream())
            return
        for (row0, span, nreal, wt1, wt2) in self.segments:
            if do1:
                _C.call('catb_pack_weights_rows', C.byref(d1), _p(wt1), _p(arena), _p(self.packed_v1), row0, span, nreal, _stream())
            if do2:
                _C.call('catb_pack_weights_rows', C.byref(d2), _p(wt2), _p(arena), _p(self.packed), row0, span, nreal, _stream())

    def _launch_timed(self, fn, reps=2, rounds=2):
        """Device time of one launch: `reps` back-to-back launches between two events, the faster of `rounds` such
        measurements (a single measurement is noisy enough to flip the choice between close candidates from run to run)."""
        fn()
        best = None
        for _ in range(rounds):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            t = e0.elapsed_time(e1) / reps
            best = t if best is None or t < best else best
        return best

    def fprop(self, x, y, bias=None, act=0, accumulate=False, y_is_f32=False, force_v1=False, stats=None):
        """stats=(sums, C, coff, per_sample): ask the epilogue to add the per-channel sum / sum of squares of the stored
        values to sums[g, 0 / 1, coff + channel] (the statistics pass of the norm layer behind the conv).  Returns True
        when the launched kernel did so (the halo kernels v2 / v3; not the gather-per-tap kernel v1)."""
        d1 = self.desc(act, accumulate, y_is_f32)
        d2 = self.desc(act, accumulate, y_is_f32, n_units=len(self.f_units))
        st = [None]    # the timing launches of the autotune run without statistics

        def v1():
            _C.call('catb_igemm_fprop', C.byref(d1), _p(self.gt), _p(x), _p(self.packed_v1), _p(bias), _p(y), _stream())

        def v2():
            es = None
            if st[0] is not None:
                es = _C.EpilogueStats()
                es.sums, es.C, es.coff, es.per_sample = st[0][0].data_ptr(), int(st[0][1]), int(st[0][2]), int(bool(st[0][3]))
                es = C.byref(es)
            if self.h_mode:   # v3: persistent pipeline, halo staged by cp.async producers (mode 1), TMA (2), TMA + reflection fringe (3)
                _C.call('catb_igemm_halo_fprop_persist', C.byref(d2), C.byref(self.hdesc), _p(self.h_steps), _p(self.h_chunks),
                        _p(x), _p(self.packed), _p(bias), _p(y), (0, 0, 1, 2)[self.h_mode], self.c_visible, es, _stream())
                return
            _C.call('catb_igemm_halo_fprop', C.byref(d2), C.byref(self.hdesc), _p(self.h_steps), _p(self.h_chunks),
                    _p(x), _p(self.packed), _p(bias), _p(y), es, _stream())

        if AUTOTUNE and self.halo is not None and self.choice is None and not force_v1 and not accumulate \
                and not torch.cuda.is_current_stream_capturing():
            # both kernels compute the same GEMM: time them once on the real operands of the first call
            # (idempotent, the output is simply rewritten) and keep the faster one
            t2, best = None, None
            for cand in self.tilings:          # every halo tiling that fits (strip width x sub-tile count)
                self._use_tiling(cand)
                tc = self._launch_timed(v2)
                if t2 is None or tc < t2:
                    t2, best = tc, cand
            self._use_tiling(best)
            t1 = self._launch_timed(v1)
            self.choice = 'v2' if t2 <= t1 else 'v1'
            self.tuned_ms = (t1, t2)
            if self.choice == 'v2':
                self.packed_v1 = None
            else:
                self.packed = None
        if self.halo is not None and not force_v1 and self.choice != 'v1':
            st[0] = stats if (FUSE_STATS and not accumulate and not y_is_f32) else None
            v2()
            return st[0] is not None
        v1()
        return False

    def _wgrad_plan(self):
        """Halo plan of the weight-gradient direction (built lazily: not every Gemm computes one)."""
        if self.w_ready:
            return
        self.w_ready = True
        self.w_halo = None
        if not USE_HALO:
            return
        lib = _C.load()
        plan = make_halo_plan(self.geo, self.units)
        if plan is None:
            return
        plan.m_sub = 1
        for tw in [self.geo.OWs] + [t for t in (64, 32, 16) if t < self.geo.OWs]:
            plan.TW = tw
            if lib.catb_igemm_halo_wgrad_fits(len(plan.planes), plan.Lh):
                break
        else:
            return
        dev = self.gt.device
        hd = _C.HaloDesc()
        hd.n_steps, hd.n_chunks, hd.n_planes = len(plan.steps), len(plan.chunks), len(plan.planes)
        for i, (pa, pb, y0, x0) in enumerate(plan.planes):
            hd.plane_pa[i], hd.plane_pb[i], hd.plane_y0[i], hd.plane_x0[i] = pa, pb, y0, x0
        hd.mul, hd.TW, hd.n_strips, hd.Wf, hd.Lh = plan.mul, plan.TW, plan.n_strips, plan.Wf, plan.Lh
        hd.Ymax, hd.Xmax, hd.m_sub = plan.Ymax, plan.Xmax, 1
        st = np.array([[pl * plan.Lh + dy * plan.Wf + dx, ci] for (ci, pl, dy, dx) in plan.steps], dtype=np.int32)
        groups = []
        for ci, (cu0, nu, first, ns) in enumerate(plan.chunks):   # groups of <= 8 taps: 8 x 64 TMEM columns
            for g0 in range(0, ns, 8):
                groups.append([ci, first + g0, min(8, ns - g0), 0])
        self.w_halo = plan
        self.w_hdesc = hd
        self.w_steps = torch.from_numpy(st).to(dev)
        self.w_chunks = torch.from_numpy(np.array(plan.chunks, dtype=np.int32)).to(dev)
        self.w_groups = torch.from_numpy(np.array(groups, dtype=np.int32)).to(dev)
        self.w_ngroups = len(groups)
        self.w_wt = units_to_device(plan.units, dev)[1]
        self.w_nunits = len(plan.units)
        # TMA-staged variant (both operands as tensor-map boxes issued by one thread): zero padding or no border taps, one
        # strip, dense output lattice
        g = self.geo
        no_border = plan.Ymax == 0 and plan.Xmax == 0 and all(pl[2] == 0 and pl[3] == 0 for pl in plan.planes)
        self.w_c_visible = 8 * max(cu0 + nu for (cu0, nu, _, _) in plan.chunks)
        self.w_tma_ok = bool(USE_TMA and (g.pad_mode != _C.PAD_REFLECT or no_border) and plan.n_strips == 1 and g.o_step == 1
                             and g.o_ph == 0 and g.o_pw == 0 and g.OHs == g.OH and g.OWs == g.OW
                             and lib.catb_igemm_halo_wgrad_tma_fits(C.byref(hd)))
        self.w_tma = False
        if self.seg_raw is not None:
            self.w_seg_wt = [units_to_device(make_halo_plan(self.geo, su).units, dev)[1] for (_r0, _sp, _nr, su) in self.seg_raw]

    def _ws_for(self, kind, d):
        """Workspace of the two-stage weight gradient: [splits][n_rows][ws_k] fp32, owned by this Gemm (its contents only
        live between the two launches of one wgrad call)."""
        key = '_ws_' + kind
        if getattr(self, key, None) is None:
            splits, ws_k = C.c_int(0), C.c_int(0)
            if kind == 'v1':
                _C.check(_C.load().catb_igemm_wgrad_ws_shape(C.byref(d), C.byref(splits), C.byref(ws_k)), 'ws_shape')
            else:
                _C.check(_C.load().catb_igemm_halo_wgrad_ws_shape(C.byref(d), C.byref(self.w_hdesc), self.w_ngroups,
                                                                  int(kind == 'v2t'), C.byref(splits), C.byref(ws_k)), 'halo ws_shape')
            ws = torch.empty(splits.value * self.n_rows * ws_k.value, dtype=torch.float32, device=self.gt.device)
            setattr(self, key, (ws, splits.value, ws_k.value))
        return getattr(self, key)

    def wgrad(self, x, y, grad_arena, force_v1=False, atomic=False):
        """grad_arena[w] += sum_rows y[row, c] * gather(x)[row, k].  Default: the deterministic two-stage form (partial
        tiles per row split into a workspace with plain stores, then catb_wgrad_unpack); atomic=True keeps the
        one-launch form with fp32 atomics from the accumulators (tests compare the two)."""
        self._wgrad_plan()
        d = self.desc()
        assert not (atomic and self.seg_raw is not None), 'N-concatenated weight gradients only exist in the two-stage form'
        queue = [None]      # second stages go to the backward pass's UnpackQueue, except while the two kernels are being timed

        def second(ws, splits, ws_k, row0, n_rows, n_units, wt, g):
            if queue[0] is not None:
                queue[0].add(ws, splits, self.n_rows, ws_k, row0, n_rows, n_units, wt, g)
            else:
                _C.call('catb_wgrad_unpack', _p(ws), splits, self.n_rows, ws_k, row0, n_rows, n_units, _p(wt), _p(g), _stream())

        def v1(g):
            if atomic:
                _C.call('catb_igemm_wgrad', C.byref(d), _p(self.gt), _p(self.wt), _p(x), _p(y), _p(g), _stream())
                return
            ws, splits, ws_k = self._ws_for('v1', d)
            _C.call('catb_igemm_wgrad_ws', C.byref(d), _p(self.gt), _p(x), _p(y), _p(ws), _stream())
            if self.seg_raw is None:
                second(ws, splits, ws_k, 0, self.n_rows, self.n_units, self.wt, g)
            else:
                if getattr(self, 'v1_seg_wt', None) is None:
                    self.v1_seg_wt = [units_to_device(su, self.gt.device)[1] for (_r0, _sp, _nr, su) in self.seg_raw]
                for (row0, _sp, nreal, _su), wt in zip(self.seg_raw, self.v1_seg_wt):
                    second(ws, splits, ws_k, row0, nreal, self.n_units, wt, g)

        def v2(g):
            if atomic:
                _C.call('catb_igemm_halo_wgrad', C.byref(d), C.byref(self.w_hdesc), _p(self.w_steps), _p(self.w_chunks),
                        _p(self.w_groups), self.w_ngroups, _p(self.w_wt), _p(x), _p(y), _p(g), _stream())
                return
            ws, splits, ws_k = self._ws_for('v2t' if self.w_tma else 'v2', d)
            _C.call('catb_igemm_halo_wgrad_ws', C.byref(d), C.byref(self.w_hdesc), _p(self.w_steps), _p(self.w_chunks),
                    _p(self.w_groups), self.w_ngroups, _p(x), _p(y), _p(ws), int(self.w_tma), self.w_c_visible, _stream())
            if self.seg_raw is None:
                second(ws, splits, ws_k, 0, self.n_rows, self.w_nunits, self.w_wt, g)
            else:
                for (row0, _sp, nreal, _su), wt in zip(self.seg_raw, self.w_seg_wt):
                    second(ws, splits, ws_k, row0, nreal, self.w_nunits, wt, g)

        if AUTOTUNE and self.w_halo is not None and self.w_choice is None and not force_v1 \
                and not torch.cuda.is_current_stream_capturing():
            # the result is accumulated into the arena, so the two kernels are timed on a scratch copy of it
            scratch = _scratch_like(grad_arena)
            t2 = self._launch_timed(lambda: v2(scratch))
            t2t = None
            if self.w_tma_ok and not atomic:
                self.w_tma = True
                t2t = self._launch_timed(lambda: v2(scratch))
                if t2t < t2:
                    t2 = t2t
                else:
                    self.w_tma = False
            t1 = self._launch_timed(lambda: v1(scratch))
            self.w_choice = 'v2' if t2 <= t1 else 'v1'
            self.w_tuned_ms = (t1, t2, t2t)
            for k in ('_ws_v1', '_ws_v2', '_ws_v2t'):      # drop the losers' workspaces
                if k != ('_ws_v1' if self.w_choice == 'v1' else ('_ws_v2t' if self.w_tma else '_ws_v2')):
                    setattr(self, k, None)
        queue[0] = _UNPACK[0]
        if self.w_halo is not None and not force_v1 and self.w_choice != 'v1':
            v2(grad_arena)
        else:
            v1(grad_arena)

    # SIMT restatements (tests only)
    def ref_fprop(self, arena, x, y, bias=None, act=0, accumulate=False, y_is_f32=False):
        d = self.desc(act, accumulate, y_is_f32)
        _C.call('catb_ref_fprop', C.byref(d), _p(self.gt), _p(self.wt), _p(arena), _p(x), _p(bias), _p(y), _stream())

    def ref_wgrad(self, x, y, grad_arena):
        d = self.desc()
        _C.call('catb_ref_wgrad', C.byref(d), _p(self.gt), _p(self.wt), _p(x), _p(y), _p(grad_arena), _stream())


# ------------------------------------------------------------------------------------------------
# element-wise / reduction wrappers
# ------------------------------------------------------------------------------------------------
def nchw_to_nhwc(src, dst: Act):
    N, Cc, H, W = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous() and (N, H, W) == (dst.N, dst.H, dst.W)
    _C.call('catb_nchw_to_nhwc', _p(src), N, Cc, H, W, _p(dst.t), dst.ld, dst.coff, _stream())


def nhwc_to_nchw(src: Act, Cc, out=None):
    if out is None:
        out = torch.empty(src.N, Cc, src.H, src.W, dtype=torch.float32, device=src.t.device)
    _C.call('catb_nhwc_to_nchw', _p(src.t), src.ld, src.coff, src.N, Cc, src.H, src.W, _p(out), _stream())
    return out


def copy_channels(src: Act, dst: Act, Cc):
    _C.call('catb_copy_channels', *src.args(), *dst.args(), src.pixels, Cc, _stream())


def norm_stats(x: Act, per_sample, sums):
    _C.call('catb_norm_stats', *x.args(), x.N, x.HW, x.C, int(per_sample), _p(sums), _stream())


def norm_finalize(sums, G, Cc, count, eps, momentum, gamma, beta, rmean, rvar, scale, shift, mean_rstd):
    _C.call('catb_norm_finalize', _p(sums), G, Cc, float(count), float(eps), float(momentum), _p(gamma), _p(beta),
            _p(rmean), _p(rvar), _p(scale), _p(shift), _p(mean_rstd), _stream())


def norm_apply(x: Act, y: Act, scale, shift, per_sample, act, residual: Act = None):
    r = residual.args() if residual is not None else (None, 0, 0)
    _C.call('catb_norm_apply', *x.args(), *y.args(), *r, x.N, x.HW, x.C, int(per_sample), _p(scale), _p(shift),
            int(act), _stream())


def norm_apply_fused(x: Act, y: Act, sums, count, eps, momentum, gamma, beta, rmean, rvar, scale, shift, mean_rstd,
                     per_sample, act, residual: Act = None):
    """norm_finalize + norm_apply in one launch (scale / shift derived per thread from the sums)."""
    r = residual.args() if residual is not None else (None, 0, 0)
    _C.call('catb_norm_apply_fused', *x.args(), *y.args(), *r, x.N, x.HW, x.C, int(per_sample), _p(sums), float(count),
            float(eps), float(momentum), _p(gamma), _p(beta), _p(rmean), _p(rvar), _p(scale), _p(shift), _p(mean_rstd),
            int(act), _stream())


def norm_bwd_reduce(dout: Act, out: Act, x: Act, per_sample, mean_rstd, act, red):
    o = out.args() if out is not None else (None, 0, 0)
    _C.call('catb_norm_bwd_reduce', *dout.args(), *o, *x.args(), x.N, x.HW, x.C, int(per_sample), _p(mean_rstd),
            int(act), _p(red), _stream())


def norm_bwd_apply(dout: Act, out: Act, x: Act, dx: Act, per_sample, mean_rstd, gamma, red, count, act, dgamma, dbeta):
    o = out.args() if out is not None else (None, 0, 0)
    _C.call('catb_norm_bwd_apply', *dout.args(), *o, *x.args(), *dx.args(), x.N, x.HW, x.C, int(per_sample),
            _p(mean_rstd), _p(gamma), _p(red), float(count), int(act), _p(dgamma), _p(dbeta), _stream())


def act_bwd(dout: Act, out: Act, dz: Act, act):
    _C.call('catb_act_bwd', *dout.args(), *out.args(), *dz.args(), dout.pixels, dout.C, int(act), _stream())


def channel_sum(x: Act, out):
    _C.call('catb_channel_sum', *x.args(), x.pixels, x.C, _p(out), _stream())


def reflect_fold(src: Act, dst: Act, p, add: Act = None):
    a = add.args() if add is not None else (None, 0, 0)
    _C.call('catb_reflect_fold', *src.args(), *dst.args(), *a, dst.N, dst.H, dst.W, dst.C, int(p), _stream())


def add(a: Act, b: Act, dst: Act):
    _C.call('catb_add', *a.args(), *b.args(), *dst.args(), a.pixels, a.C, _stream())


def dwconv_fwd(x: Act, y: Act, ksize, w_off, arena, pad_mode=_C.PAD_REFLECT):
    _C.call('catb_dwconv_fwd', *x.args(), *y.args(), x.N, x.H, x.W, x.C, _p(ksize), _p(w_off), _p(arena), int(pad_mode),
            _stream())


def dwconv_bwd_data(dy: Act, dx: Act, ksize, w_off, arena, pad_mode=_C.PAD_REFLECT):
    _C.call('catb_dwconv_bwd_data', *dy.args(), *dx.args(), dy.N, dy.H, dy.W, dy.C, _p(ksize), _p(w_off), _p(arena),
            int(pad_mode), _stream())


def dwconv_bwd_weight(x: Act, dy: Act, ksize, w_off, grad_arena, pad_mode=_C.PAD_REFLECT):
    _C.call('catb_dwconv_bwd_weight', *x.args(), *dy.args(), x.N, x.H, x.W, x.C, _p(ksize), _p(w_off),
            _p(grad_arena), int(pad_mode), _stream())


def gan_loss(pred, n, ld, mode, target_is_real, for_discriminator, grad_scale, loss, dpred: Act = None):
    d = dpred.args() if dpred is not None else (None, 0, 0)
    _C.call('catb_gan_loss', _p(pred), n, ld, _C.GAN_MODES[mode], int(target_is_real), int(for_discriminator),
            float(grad_scale), _p(loss), *d, _stream())


def recon_loss(a: Act, b: Act, Creal, kind, grad_scale, loss, da: Act = None, extra: Act = None):
    d = da.args() if da is not None else (None, 0, 0)
    e = extra.args() if extra is not None else (None, 0, 0)
    _C.call('catb_recon_loss', *a.args(), *b.args(), a.pixels, a.C, Creal, _C.RECON_KINDS[kind], float(grad_scale),
            _p(loss), *d, *e, _stream())


def gram(x: Act, G):
    _C.call('catb_gram', *x.args(), x.N, x.HW, x.C, _p(G), _stream())


def ka_finish(Gx, Gy, B, loss_scale, loss, ka_value, coef):
    _C.call('catb_ka_finish', _p(Gx), _p(Gy), B, float(loss_scale), _p(loss), _p(ka_value), _p(coef), _stream())


def ka_bwd(x: Act, coef, dx: Act, accumulate):
    _C.call('catb_ka_bwd', *x.args(), x.N, x.HW, x.C, _p(coef), *dx.args(), int(accumulate), _stream())


def adam(param, grad, m, v, lr, beta1, beta2, eps, grad_scale, step_count):
    _C.call('catb_adam', _p(param), _p(grad), _p(m), _p(v), param.numel(), _p(lr), float(beta1), float(beta2),
            float(eps), float(grad_scale), _p(step_count), _stream())


# ------------------------------------------------------------------------------------------------
# SPADE path wrappers
# ------------------------------------------------------------------------------------------------
def resize_nearest(x: Act, y: Act):
    assert x.N == y.N and x.C == y.C
    _C.call('catb_resize_nearest', *x.args(), x.H, x.W, *y.args(), y.N, y.H, y.W, y.C, _stream())


def upsample2x_bwd(dy: Act, dx: Act):
    assert dy.H == 2 * dx.H and dy.W == 2 * dx.W and dy.C == dx.C
    _C.call('catb_upsample2x_bwd', *dy.args(), *dx.args(), dx.N, dx.H, dx.W, dx.C, _stream())


def spade_modulate(x: Act, gamma: Act, beta: Act, y: Act, scale, shift, act):
    _C.call('catb_spade_modulate', *x.args(), *gamma.args(), *beta.args(), *y.args(), x.pixels, x.C, _p(scale), _p(shift),
            int(act), _stream())


def spade_modulate_bwd(dy: Act, y: Act, x: Act, gamma: Act, dgamma: Act, dbeta: Act, dn: Act, scale, shift, act):
    _C.call('catb_spade_modulate_bwd', *dy.args(), *y.args(), *x.args(), *gamma.args(), *dgamma.args(), *dbeta.args(),
            *dn.args(), x.pixels, x.C, _p(scale), _p(shift), int(act), _stream())


def act_fwd(x: Act, y: Act, act):
    _C.call('catb_act_fwd', *x.args(), *y.args(), x.pixels, x.C, int(act), _stream())


def avgpool3s2(x: Act, y: Act):
    assert y.H == (x.H + 1) // 2 and y.W == (x.W + 1) // 2 and x.C == y.C
    _C.call('catb_avgpool3s2', *x.args(), x.H, x.W, *y.args(), x.N, x.C, _stream())


def avgpool3s2_bwd(dy: Act, dx: Act, add: Act = None):
    a = add.args() if add is not None else (None, 0, 0)
    _C.call('catb_avgpool3s2_bwd', *dy.args(), *a, *dx.args(), dx.N, dx.H, dx.W, dx.C, _stream())


def maxpool2(x: Act, y: Act):
    assert y.H == x.H // 2 and y.W == x.W // 2 and x.C == y.C
    _C.call('catb_maxpool2', *x.args(), x.H, x.W, *y.args(), x.N, x.C, _stream())


def maxpool2_bwd(dy: Act, x: Act, dx: Act):
    _C.call('catb_maxpool2_bwd', *dy.args(), *x.args(), *dx.args(), x.N, x.H, x.W, x.C, _stream())


def onehot_edges(label, instance, n_label, y: Act):
    """label / instance: int32 [N,H,W] device tensors (instance may be None)."""
    assert label.dtype == torch.int32 and label.is_contiguous() and (instance is None or instance.dtype == torch.int32)
    _C.call('catb_onehot_edges', _p(label), _p(instance), y.N, y.H, y.W, int(n_label), *y.args(), y.C, _stream())


def gather_sum(arena, idx, out):
    """idx: int32 [K, n] arena offsets (-1: none); out[i] = sum_k arena[idx[k, i]]."""
    K, n = idx.shape
    assert out.numel() >= n
    _C.call('catb_gather_sum_f32', _p(arena), _p(idx), K, n, _p(out), _stream())


def scatter_add(src, idx, grad_arena):
    K, n = idx.shape
    _C.call('catb_scatter_add_f32', _p(src), _p(idx), K, n, _p(grad_arena), _stream())


def fma_vec(shift, bias, scale):
    _C.call('catb_fma_vec', _p(shift), _p(bias), _p(scale), shift.numel(), _stream())


def sn_forward(table, n, max_rows, max_cols, arena, bufs, training, tmp, sigma, w_eff):
    _C.call('catb_sn_forward', _p(table), n, max_rows, max_cols, _p(arena), _p(bufs), int(training), _p(tmp), _p(sigma),
            _p(w_eff), _stream())


def sn_backward(table, n, max_rows, max_cols, grad, w_eff, bufs, sigma, cdot):
    _C.call('catb_sn_backward', _p(table), n, max_rows, max_cols, _p(grad), _p(w_eff), _p(bufs), _p(sigma), _p(cdot),
            _stream())


# ------------------------------------------------------------------------------------------------
# x-packed 7x7 stem / head helpers
# ------------------------------------------------------------------------------------------------
def expand_x(x: Act, y: Act, Cin, taps):
    assert (x.N, x.H, x.W) == (y.N, y.H, y.W)
    _C.call('catb_expand_x', *x.args(), *y.args(), x.N, x.H, x.W, int(Cin), int(taps), y.C, _stream())


def shift_sum(P: Act, out: Act, Cout, taps, bias, act):
    assert P.W == out.W + taps - 1 and (P.N, P.H) == (out.N, out.H)
    _C.call('catb_shift_sum', *P.args(), *out.args(), out.N, out.H, out.W, int(Cout), int(taps), _p(bias), int(act), _stream())


def shift_expand(dz: Act, dP: Act, Cout, taps):
    assert dP.W == dz.W + taps - 1 and (dP.N, dP.H) == (dz.N, dz.H)
    _C.call('catb_shift_expand', *dz.args(), *dP.args(), dz.N, dz.H, dz.W, int(Cout), int(taps), _stream())


def tap_sum(P, out, H, W, OH, OW, R, S, pad, bias):
    """P: fp32 [N,H,W,ldp] (taps in channels 0 .. R*S-1), out: fp32 [N,OH,OW,ldo], channel 0 = bias + shifted tap sum."""
    _C.call('catb_tap_sum', _p(P), P.shape[-1], 0, _p(out), out.shape[-1], 0, P.shape[0], H, W, OH, OW, R, S, pad, _p(bias),
            _stream())


def tap_expand(dy: Act, dP: Act, R, S, pad):
    """dP[n,iy,ix,r*S+s] = dy[n, iy-r+pad, ix-s+pad, channel 0 of the slice] (0 outside)."""
    _C.call('catb_tap_expand', *dy.args(), *dP.args(), dy.N, dP.H, dP.W, dy.H, dy.W, R, S, pad, _stream())


class PackBatch:
    """One-launch re-packing of the GEMM weight images of a network (catb_pack_weights_batch).  The job table is
    built from the Gemm objects once their kernel choice is final (after autotuning the unused image of a GEMM is
    dropped), and rebuilt whenever that set changes."""

    def __init__(self, gemms, device):
        self.gemms, self.dev = list(gemms), device
        self.key, self.table, self.n = None, None, 0

    def _signature(self):
        return tuple((id(g), g.choice) for g in self.gemms)

    def _build(self):
        jobs = [j for g in self.gemms for j in g.pack_jobs()]
        arr = (_C.PackJob * max(len(jobs), 1))()
        big = 1
        for i, (wt, packed, n_tile, n_units, row0, span, nreal) in enumerate(jobs):
            n_chunks = (n_units + 7) // 8
            arr[i] = _C.PackJob(wt.data_ptr(), packed.data_ptr(), n_tile, n_units, n_chunks, row0, span, nreal)
            big = max(big, span * n_chunks * 8)
        raw = np.frombuffer(arr, dtype=np.uint8).copy()
        self.table = torch.from_numpy(raw).to(self.dev)
        self.n = len(jobs)
        self.blocks = max(1, min(16, (big + 255) // 256))
        self.key = self._signature()

    def run(self, arena):
        if not arena.is_cuda:            # host tensors only exist under the test emulation, which patches Gemm.pack
            for g in self.gemms:
                g.pack(arena)
            return
        if torch.cuda.is_current_stream_capturing():
            assert self.key == self._signature(), 'pack table must be built before graph capture'
        elif self.key != self._signature():
            self._build()
        if self.n:
            _C.call('catb_pack_weights_batch', _p(self.table), self.n, self.blocks, _p(arena), _stream())
