"""One CAT distillation step (teacher fwd, student fwd/bwd, discriminator update, KA / GAN / L1 losses,
two Adam updates) on libcatb200 kernels.

Restates InceptionDistiller.optimize_parameters (distillers/inception_distiller.py:179-188):
    forward (:100-104) -> backward_D (base_inception_distiller.py:293-312) -> optimizer_D.step
    -> backward_G (inception_distiller.py:159-177) -> optimizer_G.step
with the same arithmetic order of the phases.  Deviations from the reference, all stated in DESIGN.md:
bf16 activations / weights with fp32 accumulation, the two discriminator passes of backward_D are
back-propagated one after the other instead of jointly (same gradients), conv biases that feed a
normalisation layer are not updated (their gradient is analytically zero).
"""
import os

import torch

from . import ops, parallel
from .adaptors import Adaptors
from .engine import MAPPING_LAYERS, DisNet, GenNet
from .ops import Act

LOSS_NAMES = ['G_gan', 'G_distill', 'G_recon', 'D_fake', 'D_real', 'G_distill0', 'G_distill1', 'G_distill2',
              'G_distill3']


class DistillStep:
    def __init__(self, teacher_arch, student_arch, D_arch, hp, B, H, W, device='cuda:0', world_size=1,
                 use_cuda_graph=False):
        ops.require_cuda()
        self.hp, self.B, self.H, self.W, self.dev = dict(hp), B, H, W, device
        self.world_size = world_size
        self.aligned = bool(hp['aligned'])
        assert D_arch['input_nc'] == (6 if self.aligned else 3)
        # teacher_arch None: no frozen teacher and no KA terms -- the pix2pix TEACHER-TRAINING step
        # (cat_b200/train_engine.py, models/pix2pix_model.py:203-212) is this step without them
        self.T = GenNet(teacher_arch, B, H, W, device, training=False, need_grad=False) if teacher_arch is not None else None
        assert self.T is not None or (self.aligned and not hp.get('lambda_distill', 0.0))
        self.S = GenNet(student_arch, B, H, W, device, training=hp.get('student_training', True), need_grad=True)
        self.D = DisNet(D_arch, B, H, W, device)
        # --distill_G_loss_type mse (inception_distiller.py:111-133): MSE(netA_i(Sact_i), Tact_i) through the adaptor convs
        self.mse = hp.get('distill_loss_type', 'ka') == 'mse'
        assert hp.get('distill_loss_type', 'ka') in ('ka', 'mse')
        self.A = None
        if self.mse and hp.get('lambda_distill', 0.0) > 0:
            cS, cT = student_arch['widths'][2], teacher_arch['widths'][2]
            self.A = Adaptors([(self.S.acts[n], cS, self.T.acts[n], cT) for n in MAPPING_LAYERS], device)
        f32 = dict(dtype=torch.float32, device=device)
        self.real_A = torch.zeros(B, 3, H, W, **f32)
        self.real_B = torch.zeros(B, 3, H, W, **f32)
        self.xA = Act.empty(B, H, W, 3, device, zero=True)     # NHWC bf16 copies of the inputs
        self.xB = Act.empty(B, H, W, 3, device, zero=True)
        self.d_in_fake = Act.empty(B, H, W, D_arch['input_nc'], device, zero=True)
        self.d_in_real = Act.empty(B, H, W, D_arch['input_nc'], device, zero=True)
        self.dS = Act.empty(B, H, W, 3, device, zero=True)      # gradient w.r.t. the student output
        self.dS_gan = Act.empty(B, H, W, 3, device, zero=True)
        oh, ow = self.D.layers[-1].oh, self.D.layers[-1].ow
        self.dpred = Act.empty(B, oh, ow, 8, device, zero=True)
        self.losses = torch.zeros(16, **f32)                    # see LOSS_SLOTS
        self.ka_vals = torch.zeros(4, **f32)
        self.Gx = torch.zeros(4, B, B, **f32)
        self.Gy = torch.zeros(4, B, B, **f32)
        self.coef = torch.zeros(4, B, B, **f32)
        self.lr_G = torch.full((1,), float(hp['lr']), **f32)
        self.lr_D = torch.full((1,), float(hp['lr']), **f32)
        self.step_G = torch.zeros(1, dtype=torch.int32, device=device)
        self.step_D = torch.zeros(1, dtype=torch.int32, device=device)
        self.step_A = torch.zeros(1, dtype=torch.int32, device=device)
        self._graphs = None
        self.use_cuda_graph = use_cuda_graph
        # data parallel: the discriminator's gradient slices are all-reduced layer by layer during its last backward pass
        # (opt-in, CATB_EARLY_REDUCE=1: exercised over gloo on CPU only -- the one 8-GPU session of round 2 that would have
        # timed it was lost; the default keeps round 1's two all-reduces between the graph segments,
        # measured at 0.990 weak-scaling efficiency on 8 B200)
        self.early_reduce = world_size > 1 and os.environ.get('CATB_EARLY_REDUCE', '0') == '1'
        self._reducer = parallel.LayerwiseReducer(world_size)
        self.overlap_teacher = os.environ.get('CATB_NO_OVERLAP', '0') != '1'
        self._side = None

    LOSS_SLOTS = {'D_fake': 0, 'D_real': 1, 'G_gan': 2, 'G_recon': 3, 'G_distill': 4}

    # ---- state ---------------------------------------------------------------------------------
    def load(self, teacher_sd, student_sd, D_sd, netA_sds=None):
        if self.T is not None:
            self.T.load_state_dict(teacher_sd)
        self.S.load_state_dict(student_sd)
        self.D.load_state_dict(D_sd)
        if self.A is not None:
            self.A.load_state_dicts(netA_sds)

    def persistent_state(self):
        """Everything that persists between steps, by name (cat_b200/optim.py: carry_engine_state)."""
        from .optim import engine_state_from_nets
        return engine_state_from_nets({'T': self.T, 'S': self.S, 'D': self.D, 'A': self.A},
                                      {'step_G': self.step_G, 'step_D': self.step_D, 'step_A': self.step_A,
                                       'lr_G': self.lr_G, 'lr_D': self.lr_D})

    def after_state_load(self):
        for net in (self.T, self.S, self.D, self.A):
            if net is not None:
                net.pack_weights()
                for n in getattr(getattr(net, 'ns', None), 'created', []):
                    n._frozen = False          # eval-mode affine of a frozen net is recomputed from the new gamma / beta

    def set_input(self, real_A, real_B):
        """Host or device NCHW fp32 tensors -> the persistent device buffers (the H2D copy of
        BaseInceptionDistiller.set_input, base_inception_distiller.py:271-280)."""
        self.real_A.copy_(real_A, non_blocking=True)
        self.real_B.copy_(real_B, non_blocking=True)

    def set_student_training(self, training):
        """netG_student.train() / .eval(): the reference runs its first step of a run with the pruned student still in
        eval() and switches it to train() at the end of the first evaluate_model (inception_distiller.py:280).  The
        captured graphs are dropped (the launch sequence of the normalisation layers changes)."""
        if bool(training) != bool(self.S.training):
            self.S.set_training(bool(training))
            self.hp['student_training'] = bool(training)
            self._graphs = None

    def set_lr(self, lr_G, lr_D=None):
        self.lr_G.fill_(float(lr_G))
        self.lr_D.fill_(float(lr_G if lr_D is None else lr_D))

    # ---- phases --------------------------------------------------------------------------------
    def _forward_generators(self):
        """The frozen teacher's forward pass is independent of everything until backward_G needs its output and mapped
        activations, so it runs on a side stream (a parallel branch of the captured graph) next to the student forward
        and the discriminator phase; _join_teacher() closes the branch at the end of the first segment."""
        ops.nchw_to_nhwc(self.real_A, self.xA)
        ops.nchw_to_nhwc(self.real_B, self.xB)
        if self.T is None:
            pass
        elif self.overlap_teacher and self.dev != 'cpu':
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.dev)
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                self.T.forward(self.xA)
        else:
            self.T.forward(self.xA)
        self.S.forward(self.xA)

    def _join_teacher(self):
        if self.overlap_teacher and self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

    def _d_inputs(self):
        if self.aligned:
            ops.copy_channels(self.xA, self.d_in_fake, 3)
            ops.copy_channels(self.S.out, _chan_view(self.d_in_fake, 3), 3)
            ops.copy_channels(self.xA, self.d_in_real, 3)
            ops.copy_channels(self.xB, _chan_view(self.d_in_real, 3), 3)
            return self.d_in_fake, self.d_in_real
        return self.S.out, self.xB

    def _phase_D(self):
        hp, D = self.hp, self.D
        fake, real = self._d_inputs()
        D.arena.g.zero_()
        D.forward(fake)
        ops.gan_loss(D.pred, D.pred_n, 8, hp['gan_mode'], False, True, 0.5, self.losses[0:1], self.dpred)
        D.backward(self.dpred, param_grads=True, input_grad=False)
        D.forward(real)
        ops.gan_loss(D.pred, D.pred_n, 8, hp['gan_mode'], True, True, 0.5, self.losses[1:2], self.dpred)
        if self.early_reduce and self.world_size > 1:       # the gradients are final layer by layer in this (second) pass
            D.backward(self.dpred, param_grads=True, input_grad=False, grads_final_hook=self._reducer.reduce_async)
            self._reducer.join()
        else:
            D.backward(self.dpred, param_grads=True, input_grad=False)

    def _adam(self, net, lr, step):
        a = net.arena
        ops.adam(a.p, a.g, a.m, a.v, lr, self.hp['beta1'], 0.999, 1e-8, parallel.grad_scale(self.world_size), step)
        net.pack_weights()

    def _phase_G(self):
        hp, D, S, T = self.hp, self.D, self.S, self.T
        fake, _ = self._d_inputs()
        S.arena.g.zero_()
        D.forward(fake)
        ops.gan_loss(D.pred, D.pred_n, 8, hp['gan_mode'], True, False, hp['lambda_gan'], self.losses[2:3], self.dpred)
        d_in = D.backward(self.dpred, param_grads=False, input_grad=True)
        if self.aligned:
            ops.copy_channels(_chan_view(d_in, 3), self.dS_gan, 3)
            extra, target = self.dS_gan, self.xB
        else:
            extra, target = d_in, T.out
        ops.recon_loss(S.out, target, 3, hp.get('recon_loss_type', 'l1'), hp['lambda_recon'], self.losses[3:4],
                       self.dS, extra)
        act_grads = {}
        if self.A is not None:
            self.A.arena.g.zero_()
            self.ka_vals.zero_()
            scale = hp['lambda_distill'] * hp.get('ka_scale', 1.0)
            for i, n in enumerate(MAPPING_LAYERS):
                self.A.loss(i, scale, self.ka_vals[i:i + 1])
                act_grads[n] = (lambda dact, i=i: self.A.backward_into(i, dact))
        elif hp.get('lambda_distill', 0.0) > 0:
            self.Gx.zero_()
            self.Gy.zero_()
            scale = -hp['lambda_distill'] * hp.get('ka_scale', 1.0)
            for i, n in enumerate(MAPPING_LAYERS):
                ops.gram(S.acts[n], self.Gx[i])
                ops.gram(T.acts[n], self.Gy[i])
                ops.ka_finish(self.Gx[i], self.Gy[i], self.B, scale, self.losses[4:5], self.ka_vals[i:i + 1], self.coef[i])
                act_grads[n] = (lambda dact, i=i, n=n: ops.ka_bwd(S.acts[n], self.coef[i], dact, True))
        S.backward(self.dS, act_grads)

    # ---- the step ------------------------------------------------------------------------------
    def _part1(self):
        self.losses.zero_()
        self._forward_generators()
        self._phase_D()
        self._join_teacher()

    def _part2(self):
        self._adam(self.D, self.lr_D, self.step_D)
        self._phase_G()

    def _part3(self):
        self._adam(self.S, self.lr_G, self.step_G)
        if self.A is not None:          # the adaptors are the second parameter group of optimizer_G
            self._adam(self.A, self.lr_G, self.step_A)

    def step(self):
        """optimize_parameters(): three launch segments separated by the two gradient all-reduces."""
        if self.use_cuda_graph:
            if self._graphs is None:
                self._capture()
            g1, g2, g3 = self._graphs
            g1.replay()
            self._allreduce(self.D)
            g2.replay()
            self._allreduce(self.S)
            g3.replay()
        else:
            self._part1()
            self._allreduce(self.D)
            self._part2()
            self._allreduce(self.S)
            self._part3()

    def _allreduce(self, net):
        if net is self.D and self.early_reduce and self.world_size > 1:
            return                      # already reduced layer by layer inside the D phase
        parallel.reduce_gradients(net.arena.g, self.world_size)
        if net is self.S and self.A is not None:
            parallel.reduce_gradients(self.A.arena.g, self.world_size)

    def _tune_pass(self):
        """One eager step on a snapshot of every mutable tensor: each Gemm times its two forward kernels on
        its real operands (ops.Gemm.fprop) and keeps the faster; the state is then restored, so the captured
        graph starts from exactly the loaded weights."""
        state = []
        for net in (self.S, self.D, self.T, self.A):
            if net is None:
                continue
            state += [t for t in (net.arena.p, net.arena.g, net.arena.m, net.arena.v, net.bufs.p) if t is not None]
        state += [self.step_G, self.step_D, self.step_A, self.losses, self.ka_vals]
        snap = [t.clone() for t in state]
        # kernels are timed one at a time: no side-stream branches during the tuning step
        saved = (self.overlap_teacher, self.S.overlap_wgrad)
        self.overlap_teacher = self.S.overlap_wgrad = False
        self._part1()
        self._part2()
        self._part3()
        torch.cuda.synchronize()
        self.overlap_teacher, self.S.overlap_wgrad = saved
        for t, c in zip(state, snap):
            t.copy_(c)
        self.S.pack_weights()
        self.D.pack_weights()
        if self.A is not None:
            self.A.pack_weights()
        torch.cuda.synchronize()

    def _capture(self):
        self._tune_pass()
        # capture the three segments on a side stream
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        graphs = []
        from . import _C
        n0 = _C.LAUNCH_COUNT[0]
        with torch.cuda.stream(s):
            for part in (self._part1, self._part2, self._part3):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    part()
                graphs.append(g)
        torch.cuda.current_stream().wait_stream(s)
        self._graphs = graphs
        self.launches_per_step = _C.LAUNCH_COUNT[0] - n0   # libcatb200 kernels captured per step

    def get_losses(self):
        """Synchronises (float() on device scalars), like get_current_losses (base_model.py:166-188)."""
        l = self.losses.tolist()
        k = self.ka_vals.tolist()
        hp = self.hp
        scale = hp.get('ka_scale', 1.0)
        out = {'D_fake': l[0], 'D_real': l[1], 'G_gan': l[2] * hp['lambda_gan'], 'G_recon': l[3] * hp['lambda_recon'], 'G_distill': l[4]}
        if self.A is not None:          # 'mse': the slots hold the four unscaled MSE terms
            out['G_distill'] = hp['lambda_distill'] * scale * sum(k[:4])
        for i in range(4):
            out['G_distill%d' % i] = (k[i] if self.A is not None else -k[i]) * scale
        return out


def _chan_view(act: Act, c):
    """View of `act` whose slice starts at channel c (may be unaligned; only for copy_channels)."""
    v = Act(act.t)
    v.coff, v.C = act.coff + c, act.ld - act.coff - c
    return v
