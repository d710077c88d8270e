"""One CAT SPADE distillation step on libcatb200 kernels.

Restates BaseSPADEDistiller.optimize_parameters (distillers/base_spade_distiller.py:226-234):
    backward_G (models/spade_model.py:189-196 -> compute_G_loss, base_spade_distiller_modules.py:128-158)
    -> optimizer_G.step -> backward_D (-> compute_D_loss, :160-175) -> optimizer_D.step
with the reference's order of the phases: the generator is updated FIRST, and the discriminator phase runs a
second, gradient-free student forward with the updated weights.  Preprocessing (one-hot labels + instance
edges, models/spade_model.py:142-179) is part of the step.

Deviations from the reference, stated in DESIGN.md: bf16 activations / GEMM operands with fp32 accumulation;
conv biases that feed a BatchNorm are not updated (analytically zero gradient); per-rank BatchNorm statistics
under data parallelism (= the reference's single-GPU semantics on every rank).
"""
import os

import torch

from . import ops, parallel
from .adaptors import Adaptors
from .ops import Act
from .spade_engine import MAPPING_LAYERS, VGG_WEIGHTS, MultiScaleDis, SpadeGenNet, VggNet, dis_feature

LOSS_NAMES = ['G_gan', 'G_feat', 'G_vgg', 'G_distill', 'D_real', 'D_fake', 'G_distill0', 'G_distill1', 'G_distill2']


class SpadeDistillStep:
    # slots of the device loss vector
    S_GAN, S_FEAT, S_DISTILL, S_DFAKE, S_DREAL, S_VGG0 = 0, 1, 2, 3, 4, 8

    def __init__(self, teacher_arch, student_arch, D_arch, hp, B, H, W, device='cuda:0', world_size=1, use_cuda_graph=False):
        ops.require_cuda()
        self.hp, self.B, self.H, self.W, self.dev = dict(hp), B, H, W, device
        self.world_size = world_size
        snc = student_arch['semantic_nc']
        assert (teacher_arch is None or teacher_arch['semantic_nc'] == snc) and D_arch['input_nc'] == snc + 3
        self.snc, self.n_label = snc, int(hp['n_label'])
        f32 = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.label = torch.zeros(B, H, W, **i32)
        self.instance = torch.zeros(B, H, W, **i32)
        self.image = torch.zeros(B, 3, H, W, **f32)
        self.seg = Act.empty(B, H, W, snc, device, zero=True)
        self.xB = Act.empty(B, H, W, 3, device, zero=True)
        # teacher_arch None: no frozen teacher and no KA terms -- SPADEModel's TEACHER-TRAINING step
        # (cat_b200/train_engine.py, models/spade_model.py:207-215) is this step without them
        self.T = SpadeGenNet(teacher_arch, self.seg, device, training=False, need_grad=False) if teacher_arch is not None else None
        assert self.T is not None or not hp.get('lambda_distill', 0.0)
        self.S = SpadeGenNet(student_arch, self.seg, device, training=hp.get('student_training', True), need_grad=True)
        self.D = MultiScaleDis(D_arch, 2 * B, H, W, device)
        # --distill_G_loss_type mse (spade_distiller_modules.py:23-25): MSE(netA_i(Sact_i), Tact_i) through the adaptor convs
        assert hp.get('distill_loss_type', 'ka') in ('ka', 'mse')
        self.A = None
        if hp.get('distill_loss_type', 'ka') == 'mse' and hp.get('lambda_distill', 0.0) > 0:
            self.A = Adaptors([(self.S.acts[n], student_arch['blocks'][n]['fout'], self.T.acts[n], teacher_arch['blocks'][n]['fout'])
                               for n in MAPPING_LAYERS], device)
        self.step_A = torch.zeros(1, dtype=torch.int32, device=device)
        self.V = VggNet(B, H, W, device)
        cin = D_arch['input_nc']
        self.d_in = Act.empty(2 * B, H, W, cin, device, zero=True)       # [seg | fake ; seg | real]
        self.dS = Act.empty(B, H, W, 3, device, zero=True)
        self.dS_gan = Act.empty(B, H, W, 3, device, zero=True)
        self.dpreds = [Act.empty(2 * B, n.layers[-1].oh, n.layers[-1].ow, 8, device, zero=True) for n in self.D.nets]
        self.losses = torch.zeros(16, **f32)
        nl = len(MAPPING_LAYERS)
        self.ka_vals = torch.zeros(nl, **f32)
        self.Gx = torch.zeros(nl, B, B, **f32)
        self.Gy = torch.zeros(nl, B, B, **f32)
        self.coef = torch.zeros(nl, B, B, **f32)
        self.lr_G = torch.full((1,), float(hp['lr_G']), **f32)
        self.lr_D = torch.full((1,), float(hp['lr_D']), **f32)
        self.step_G = torch.zeros(1, dtype=torch.int32, device=device)
        self.step_D = torch.zeros(1, dtype=torch.int32, device=device)
        self._graphs = None
        self.use_cuda_graph = use_cuda_graph
        self.overlap = os.environ.get('CATB_NO_OVERLAP', '0') != '1'
        self._side = None

    # ---- state ---------------------------------------------------------------------------------
    def load(self, teacher_sd, student_sd, D_sd, vgg_sd, netA_sds=None):
        if self.T is not None:
            self.T.load_state_dict(teacher_sd)
        self.S.load_state_dict(student_sd)
        self.D.load_state_dict(D_sd)
        self.V.load_state_dict(vgg_sd)
        if self.A is not None:
            self.A.load_state_dicts(netA_sds)

    def persistent_state(self):
        """Everything that persists between steps, by name (cat_b200/optim.py: carry_engine_state)."""
        from .optim import engine_state_from_nets
        return engine_state_from_nets({'T': self.T, 'S': self.S, 'D': self.D, 'V': self.V, 'A': self.A},
                                      {'step_G': self.step_G, 'step_D': self.step_D, 'step_A': self.step_A,
                                       'lr_G': self.lr_G, 'lr_D': self.lr_D})

    def after_state_load(self):
        for net in (self.T, self.S, self.V, self.A):
            if net is not None:
                net.pack_weights()
                for n in getattr(net, 'norms', []):
                    n._frozen = False          # eval-mode affine of a frozen net is recomputed from the new gamma / beta
        self.D.spectral_forward(training=False)     # effective weights + GEMM images from weight_orig / u / v

    def set_input(self, label, instance, image):
        """label / instance: [B,1,H,W] (any integer or float dtype, host or device), image: [B,3,H,W] fp32 in
        [-1,1] -- the dict entries of SPADEModel.set_input (models/spade_model.py:132-136)."""
        B, H, W = self.B, self.H, self.W
        self.label.copy_(label.reshape(B, H, W).to(torch.int32), non_blocking=True)
        self.instance.copy_(instance.reshape(B, H, W).to(torch.int32), non_blocking=True)
        self.image.copy_(image, non_blocking=True)

    def set_student_training(self, training):
        """netG_student.train() / .eval(): like the Inception distiller, the reference runs its first step of a run with the
        student still in eval() (model_profiling in BaseSPADEDistiller.setup, base_spade_distiller.py:178-190) and switches
        it to train() at the end of the first evaluate_model (spade_distiller.py:170).  Captured graphs are dropped."""
        if bool(training) != bool(self.S.training):
            self.S.set_training(bool(training))
            self.hp['student_training'] = bool(training)
            self._graphs = None

    def set_lr(self, lr_G, lr_D):
        self.lr_G.fill_(float(lr_G))
        self.lr_D.fill_(float(lr_D))

    # ---- phases --------------------------------------------------------------------------------
    def _preprocess(self):
        ops.onehot_edges(self.label, self.instance, self.n_label, self.seg)
        ops.nchw_to_nhwc(self.image, self.xB)

    def _d_input(self, fake: Act):
        """fake_and_real = cat([cat(seg, fake), cat(seg, real)], dim=0) (spade_model_modules.py:136-141)."""
        B, snc = self.B, self.snc
        lo, hi = Act(self.d_in.t[:B]), Act(self.d_in.t[B:])
        ops.copy_channels(self.seg, lo, snc)
        ops.copy_channels(fake, _chan_view(lo, snc), 3)
        ops.copy_channels(self.seg, hi, snc)
        ops.copy_channels(self.xB, _chan_view(hi, snc), 3)
        return self.d_in

    def _adam(self, net, lr, step):
        a, hp = net.arena, self.hp
        ops.adam(a.p, a.g, a.m, a.v, lr, hp['beta1'], hp['beta2'], 1e-8, parallel.grad_scale(self.world_size), step)
        net.pack_weights()

    def _phase_G(self):
        hp, D, S, T, V, B = self.hp, self.D, self.S, self.T, self.V, self.B
        S.arena.g.zero_()
        # two independent branches next to the student / discriminator work (parallel branches of the captured graph):
        # the frozen teacher (needed by the KA terms) and the VGG features of the real image (needed by the VGG loss)
        main = torch.cuda.current_stream() if self.overlap and self.dev != 'cpu' else None
        if main is not None:
            if self._side is None:
                self._side = (torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev))
            for st, fn in ((self._side[0], T.forward if T is not None else (lambda: None)),
                           (self._side[1], lambda: V.forward(self.xB, save_ref=True))):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    fn()
        elif T is not None:
            T.forward()
        S.forward()
        nets = D.forward(self._d_input(S.out))
        num_D = len(nets)
        # GAN loss on the fake half of every scale (hinge, generator side); the real half gets zero gradient
        for i, net in enumerate(nets):
            n = B * net.layers[-1].oh * net.layers[-1].ow
            self.dpreds[i].t.zero_()
            ops.gan_loss(net.pred[:B], n, 8, 'hinge', True, False, hp['lambda_gan'] / num_D, self.losses[self.S_GAN:self.S_GAN + 1],
                         Act(self.dpreds[i].t[:B]))
        # feature matching: L1 between the intermediate outputs of the fake and the (detached) real half
        fscale = hp['lambda_feat'] / num_D

        def feat_hook(i, li, d):
            f = dis_feature(nets[i], li)
            fake, real, dfake = Act(f.t[:B]), Act(f.t[B:]), Act(d.t[:B])
            ops.recon_loss(fake, real, nets[i].layers[li].cout, 'l1', fscale, self.losses[self.S_FEAT:self.S_FEAT + 1], dfake, dfake)

        d_in = D.backward(self.dpreds, param_grads=False, input_grad=True, act_grad_hook=feat_hook)
        ops.copy_channels(_chan_view(Act(d_in.t[:B]), self.snc), self.dS_gan, 3)
        # VGG loss: features of the real image first (kept at the five taps), then the fake image with backward
        if main is not None:
            main.wait_stream(self._side[1])
        else:
            V.forward(self.xB, save_ref=True)
        V.forward(S.out)
        d_vgg = V.loss_and_backward(self.losses[self.S_VGG0:self.S_VGG0 + 5], hp['lambda_vgg'])
        ops.add(self.dS_gan, d_vgg, self.dS)
        act_grads = {}
        if main is not None:
            main.wait_stream(self._side[0])
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
                ops.ka_finish(self.Gx[i], self.Gy[i], B, scale, self.losses[self.S_DISTILL:self.S_DISTILL + 1], self.ka_vals[i:i + 1],
                              self.coef[i])
                act_grads[n] = (lambda dact, i=i, n=n: ops.ka_bwd(S.acts[n], self.coef[i], dact, True))
        S.backward(self.dS, act_grads)

    def _phase_D(self):
        hp, D, S, B = self.hp, self.D, self.S, self.B
        D.arena.g.zero_()
        S.forward()                       # no-grad forward with the updated student (train mode: BN statistics move)
        nets = D.forward(self._d_input(S.out))
        num_D = len(nets)
        for i, net in enumerate(nets):
            n = B * net.layers[-1].oh * net.layers[-1].ow
            ops.gan_loss(net.pred[:B], n, 8, 'hinge', False, True, 1.0 / num_D, self.losses[self.S_DFAKE:self.S_DFAKE + 1],
                         Act(self.dpreds[i].t[:B]))
            ops.gan_loss(net.pred[B:], n, 8, 'hinge', True, True, 1.0 / num_D, self.losses[self.S_DREAL:self.S_DREAL + 1],
                         Act(self.dpreds[i].t[B:]))
        D.backward(self.dpreds, param_grads=True, input_grad=False)
        D.finish_param_grads()

    def _allreduce(self, net):
        parallel.reduce_gradients(net.arena.g, self.world_size)
        if net is self.S and self.A is not None:
            parallel.reduce_gradients(self.A.arena.g, self.world_size)

    # ---- the step ------------------------------------------------------------------------------
    def _part1(self):
        self.losses.zero_()
        self._preprocess()
        self._phase_G()

    def _part2(self):
        self._adam(self.S, self.lr_G, self.step_G)
        if self.A is not None:          # the adaptors are parameters of optimizer_G
            self._adam(self.A, self.lr_G, self.step_A)
        self._phase_D()

    def _part3(self):
        self._adam(self.D, self.lr_D, self.step_D)

    def step(self):
        """optimize_parameters(): three launch segments separated by the two gradient all-reduces."""
        if self.use_cuda_graph:
            if self._graphs is None:
                self._capture()
            g1, g2, g3 = self._graphs
            g1.replay()
            self._allreduce(self.S)
            g2.replay()
            self._allreduce(self.D)
            g3.replay()
        else:
            self._part1()
            self._allreduce(self.S)
            self._part2()
            self._allreduce(self.D)
            self._part3()

    def _mutable_state(self):
        state = []
        for net in (self.S, self.D, self.T, self.A):
            if net is None:
                continue
            state += [t for t in (net.arena.p, net.arena.g, net.arena.m, net.arena.v, net.bufs.p) if t is not None]
        return state + [self.step_G, self.step_D, self.step_A, self.losses, self.ka_vals]

    def _tune_pass(self):
        """One eager step on a snapshot of every mutable tensor (each Gemm autotunes on its real operands), then the
        state is restored so that the captured graphs start from exactly the loaded weights."""
        state = self._mutable_state()
        snap = [t.clone() for t in state]
        saved, self.overlap = (self.overlap, self.S.overlap_wgrad), False     # kernels are timed one at a time while tuning
        self.S.overlap_wgrad = False
        self._part1()
        self._part2()
        self._part3()
        torch.cuda.synchronize()
        self.overlap, self.S.overlap_wgrad = saved
        for t, c in zip(state, snap):
            t.copy_(c)
        self.S.pack_weights()
        if self.A is not None:
            self.A.pack_weights()
        self.D.spectral_forward(training=False)
        torch.cuda.synchronize()

    def _capture(self):
        from . import _C
        self._tune_pass()
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        graphs = []
        n0 = _C.LAUNCH_COUNT[0]
        with torch.cuda.stream(s):
            for part in (self._part1, self._part2, self._part3):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    part()
                graphs.append(g)
        torch.cuda.current_stream().wait_stream(s)
        self._graphs = graphs
        self.launches_per_step = _C.LAUNCH_COUNT[0] - n0

    def get_losses(self):
        """Synchronises, like get_current_losses (models/base_model.py:166-188).  Values as the reference reports
        them (spade_model.py:189-205): every G term already multiplied by its lambda."""
        l, k, hp = self.losses.tolist(), self.ka_vals.tolist(), self.hp
        num_D = len(self.D.nets)
        out = {'G_gan': l[self.S_GAN] * hp['lambda_gan'] / num_D, 'G_feat': l[self.S_FEAT] * hp['lambda_feat'] / num_D,
               'G_vgg': sum(w * v for w, v in zip(VGG_WEIGHTS, l[self.S_VGG0:self.S_VGG0 + 5])) * hp['lambda_vgg'],
               'G_distill': l[self.S_DISTILL], 'D_fake': l[self.S_DFAKE] / num_D, 'D_real': l[self.S_DREAL] / num_D}
        scale = hp.get('ka_scale', 1.0)
        if self.A is not None:          # 'mse': the slots hold the unscaled MSE terms
            out['G_distill'] = hp['lambda_distill'] * scale * sum(k[:len(MAPPING_LAYERS)])
        for i in range(len(MAPPING_LAYERS)):
            out['G_distill%d' % i] = (k[i] if self.A is not None else -k[i]) * scale
        return out


def _chan_view(act: Act, c):
    """View of `act` whose slice starts at channel c (may be unaligned; only for copy_channels)."""
    v = Act(act.t)
    v.coff, v.C = act.coff + c, act.ld - act.coff - c
    return v
