"""TEACHER-TRAINING steps on libcatb200 kernels (the step on the other side of the teacher
checkpoint).  Same networks, kernels and execution model as the distillation steps (engine.py, distill_engine.py,
spade_distill_engine.py): static launch sequences with hand-derived backward passes, flat fp32 arenas, CUDA graphs.

  Pix2PixTrainStep  -- Pix2PixModel.optimize_parameters  (models/pix2pix_model.py:203-212)
  CycleGANTrainStep -- CycleGANModel.optimize_parameters (models/cycle_gan_model.py:292-303), incl. the ImagePool
                       history buffers (utils/image_pool.py) kept on the device
  SpadeTrainStep    -- SPADEModel.optimize_parameters    (models/spade_model.py:207-215)
"""
import random

import torch

from . import ops, parallel
from .distill_engine import DistillStep
from .engine import DisNet, GenNet
from .ops import Act
from .spade_distill_engine import SpadeDistillStep


class Pix2PixTrainStep(DistillStep):
    """forward (pix2pix_model.py:153-155) -> backward_D (:157-172) -> optimizer_D.step -> backward_G (:174-201: GAN *
    lambda_gan + recon * lambda_recon; the comp-cost term is off at its default weight 0) -> optimizer_G.step: the
    distillation step of distill_engine.DistillStep without the frozen teacher and the KA terms, on aligned pairs."""

    LOSS_NAMES = ['G_gan', 'G_recon', 'D_real', 'D_fake']

    def __init__(self, G_arch, D_arch, hp, B, H, W, device='cuda:0', world_size=1, use_cuda_graph=False):
        hp = dict(hp, aligned=True, lambda_distill=0.0, student_training=True)
        super().__init__(None, G_arch, D_arch, hp, B, H, W, device=device, world_size=world_size, use_cuda_graph=use_cuda_graph)
        self.G = self.S

    def load(self, G_sd, D_sd):
        super().load(None, G_sd, D_sd)

    def get_losses(self):
        l = super().get_losses()
        return {k: l[k] for k in self.LOSS_NAMES}


class DeviceImagePool:
    """utils/image_pool.py:5-53 with the history kept on the device in the discriminator's input layout (NHWC, channels
    padded to the 8-channel unit).  The decisions are drawn on the host from Python's global ``random`` in the
    reference's order (one uniform(0,1) per image once the pool is full, then randint(0, pool_size-1)), so a run seeded
    like the reference makes the same choices.  The copies are plain device-to-device tensor copies on the current
    stream, outside the captured graphs (their source / destination slots change from step to step)."""

    def __init__(self, pool_size, like: Act):
        self.pool_size, self.num = int(pool_size), 0
        if self.pool_size > 0:
            self.images = torch.zeros((self.pool_size,) + tuple(like.t.shape[1:]), dtype=like.t.dtype, device=like.t.device)
            self.out = Act(torch.zeros_like(like.t))

    def query(self, fake: Act) -> Act:
        if self.pool_size == 0:
            return fake
        src, out = fake.t, self.out.t
        for i in range(src.shape[0]):
            if self.num < self.pool_size:
                self.images[self.num].copy_(src[i])
                self.num += 1
                out[i].copy_(src[i])
            elif random.uniform(0, 1) > 0.5:
                j = random.randint(0, self.pool_size - 1)
                out[i].copy_(self.images[j])
                self.images[j].copy_(src[i])
            else:
                out[i].copy_(src[i])
        return self.out


class CycleGANTrainStep:
    """forward (cycle_gan_model.py:221-226) -> backward_G (:260-290, discriminators frozen) -> optimizer_G.step over both
    generators -> backward_D_A / backward_D_B (:228-258, real first, then the pooled fake) -> optimizer_D.step.

    Every generator is applied three times per step (to the real image of its source domain, to the other generator's
    output for the cycle, to the real image of its target domain for the identity term): three GenNet compilations that
    share one parameter / gradient / running-statistics arena, run in the reference's call order.  The cycle term
    back-propagates through the INPUT of the second application (GenNet input_grad) into the first one, where it meets
    the GAN gradient coming back through the frozen discriminator."""

    LOSS_NAMES = ['D_A', 'G_A', 'G_cycle_A', 'G_idt_A', 'D_B', 'G_B', 'G_cycle_B', 'G_idt_B']
    SLOT = {n: i for i, n in enumerate(LOSS_NAMES)}

    def __init__(self, G_arch, D_arch, hp, B, H, W, device='cuda:0', world_size=1, use_cuda_graph=False):
        ops.require_cuda()
        self.hp, self.B, self.H, self.W, self.dev = dict(hp), B, H, W, device
        self.world_size, self.use_cuda_graph = world_size, use_cuda_graph
        assert G_arch['input_nc'] == G_arch['output_nc'] == D_arch['input_nc'] == 3
        self.idt = hp['lambda_identity'] > 0
        mk = lambda share, ig: GenNet(G_arch, B, H, W, device, training=True, need_grad=True, share=share, input_grad=ig)
        # G_A: A -> B, G_B: B -> A;  *_real on the real source image, *_cyc on the other generator's output, *_idt on the
        # real target image
        self.GA_real = mk(None, False)
        self.GA_cyc = mk(self.GA_real, True)
        self.GB_real = mk(None, False)
        self.GB_cyc = mk(self.GB_real, True)
        self.GA_idt = mk(self.GA_real, False) if self.idt else None
        self.GB_idt = mk(self.GB_real, False) if self.idt else None
        self.G_A, self.G_B = self.GA_real, self.GB_real            # owners of the arenas
        self.D_A = DisNet(D_arch, B, H, W, device)                  # G_A(A) vs. B
        self.D_B = DisNet(D_arch, B, H, W, device)                  # G_B(B) vs. A
        f32 = dict(dtype=torch.float32, device=device)
        self.real_A = torch.zeros(B, 3, H, W, **f32)
        self.real_B = torch.zeros(B, 3, H, W, **f32)
        act = lambda: Act.empty(B, H, W, 3, device, zero=True)
        self.xA, self.xB = act(), act()
        self.d_rec_A, self.d_rec_B, self.d_idt_A, self.d_idt_B = act(), act(), act(), act()
        self.d_fake_A, self.d_fake_B = act(), act()
        oh, ow = self.D_A.layers[-1].oh, self.D_A.layers[-1].ow
        self.dpred = Act.empty(B, oh, ow, 8, device, zero=True)
        self.losses = torch.zeros(16, **f32)
        self.lr_G = torch.full((1,), float(hp['lr']), **f32)
        self.lr_D = torch.full((1,), float(hp['lr']), **f32)
        self.step_GA = torch.zeros(1, dtype=torch.int32, device=device)
        self.step_GB = torch.zeros(1, dtype=torch.int32, device=device)
        self.step_DA = torch.zeros(1, dtype=torch.int32, device=device)
        self.step_DB = torch.zeros(1, dtype=torch.int32, device=device)
        self.pool_A = DeviceImagePool(hp.get('pool_size', 50), self.GB_real.out)     # history of fake_A
        self.pool_B = DeviceImagePool(hp.get('pool_size', 50), self.GA_real.out)     # history of fake_B
        self.d_in_fake_B, self.d_in_fake_A = self.GA_real.out, self.GB_real.out      # what the D phase reads (set per step)
        self._graphs = None

    # ---- state ---------------------------------------------------------------------------------
    def _gens(self, which):
        return [g for g in ((self.GA_real, self.GA_cyc, self.GA_idt) if which == 'A' else
                            (self.GB_real, self.GB_cyc, self.GB_idt)) if g is not None]

    def load(self, G_A_sd, G_B_sd, D_A_sd, D_B_sd):
        self.G_A.load_state_dict(G_A_sd)
        self.G_B.load_state_dict(G_B_sd)
        self._pack_generators()
        self.D_A.load_state_dict(D_A_sd)
        self.D_B.load_state_dict(D_B_sd)

    def persistent_state(self):
        """Everything that persists between steps, by name (cat_b200/optim.py: carry_engine_state)."""
        from .optim import engine_state_from_nets
        extra = {'step_GA': self.step_GA, 'step_GB': self.step_GB, 'step_DA': self.step_DA, 'step_DB': self.step_DB,
                 'lr_G': self.lr_G, 'lr_D': self.lr_D}
        for name in ('pool_A', 'pool_B'):
            pool = getattr(self, name)
            if pool.pool_size > 0:
                extra[name + '.images'] = pool.images
        return engine_state_from_nets({'G_A': self.G_A, 'G_B': self.G_B, 'D_A': self.D_A, 'D_B': self.D_B}, extra)

    def after_state_load(self):
        self._pack_generators()
        self.D_A.pack_weights()
        self.D_B.pack_weights()
        for g in self._gens('A') + self._gens('B'):
            for n in g.ns.created:
                n._frozen = False

    def _pack_generators(self):
        for w in 'AB':
            for g in self._gens(w):
                g.pack_weights()

    def set_input(self, real_A, real_B):
        self.real_A.copy_(real_A, non_blocking=True)
        self.real_B.copy_(real_B, non_blocking=True)

    def set_lr(self, lr_G, lr_D=None):
        self.lr_G.fill_(float(lr_G))
        self.lr_D.fill_(float(lr_G if lr_D is None else lr_D))

    # ---- phases --------------------------------------------------------------------------------
    def _forward(self):
        ops.nchw_to_nhwc(self.real_A, self.xA)
        ops.nchw_to_nhwc(self.real_B, self.xB)
        self.fake_B = self.GA_real.forward(self.xA)
        self.rec_A = self.GB_cyc.forward(self.fake_B)
        self.fake_A = self.GB_real.forward(self.xB)
        self.rec_B = self.GA_cyc.forward(self.fake_A)

    def _gan_through(self, D, fake, slot):
        """criterionGAN(netD(fake), True) with the discriminator frozen; returns d loss / d fake."""
        D.forward(fake)
        ops.gan_loss(D.pred, D.pred_n, 8, self.hp['gan_mode'], True, True, 1.0, self.losses[slot:slot + 1], self.dpred)
        return D.backward(self.dpred, param_grads=False, input_grad=True)

    def _phase_G(self):
        hp, S = self.hp, self.SLOT
        lA, lB, lI = hp['lambda_A'], hp['lambda_B'], hp['lambda_identity']
        self.G_A.arena.g.zero_()
        self.G_B.arena.g.zero_()
        if self.idt:
            idt_A = self.GA_idt.forward(self.xB)
            ops.recon_loss(idt_A, self.xB, 3, 'l1', lB * lI, self.losses[S['G_idt_A']:S['G_idt_A'] + 1], self.d_idt_A)
            idt_B = self.GB_idt.forward(self.xA)
            ops.recon_loss(idt_B, self.xA, 3, 'l1', lA * lI, self.losses[S['G_idt_B']:S['G_idt_B'] + 1], self.d_idt_B)
        # cycle A: L1(G_B(G_A(A)), A) -> through G_B's input into fake_B, where the GAN gradient from D_A is added
        ops.recon_loss(self.rec_A, self.xA, 3, 'l1', lA, self.losses[S['G_cycle_A']:S['G_cycle_A'] + 1], self.d_rec_A)
        d_cyc = self.GB_cyc.backward(self.d_rec_A)
        d_gan = self._gan_through(self.D_A, self.fake_B, S['G_A'])
        ops.add(d_gan, d_cyc, self.d_fake_B)
        self.GA_real.backward(self.d_fake_B)
        # cycle B
        ops.recon_loss(self.rec_B, self.xB, 3, 'l1', lB, self.losses[S['G_cycle_B']:S['G_cycle_B'] + 1], self.d_rec_B)
        d_cyc = self.GA_cyc.backward(self.d_rec_B)
        d_gan = self._gan_through(self.D_B, self.fake_A, S['G_B'])
        ops.add(d_gan, d_cyc, self.d_fake_A)
        self.GB_real.backward(self.d_fake_A)
        if self.idt:
            self.GA_idt.backward(self.d_idt_A)
            self.GB_idt.backward(self.d_idt_B)

    def _backward_D(self, D, real, fake, slot):
        """backward_D_basic (cycle_gan_model.py:228-246): real first, then the (pooled) fake; (real + fake) * 0.5."""
        mode = self.hp['gan_mode']
        D.arena.g.zero_()
        D.forward(real)
        ops.gan_loss(D.pred, D.pred_n, 8, mode, True, True, 0.5, self.losses[slot:slot + 1], self.dpred)
        D.backward(self.dpred, param_grads=True, input_grad=False)
        D.forward(fake)
        ops.gan_loss(D.pred, D.pred_n, 8, mode, False, True, 0.5, self.losses[slot:slot + 1], self.dpred)
        D.backward(self.dpred, param_grads=True, input_grad=False)

    def _phase_D(self):
        self._backward_D(self.D_A, self.xB, self.d_in_fake_B, self.SLOT['D_A'])
        self._backward_D(self.D_B, self.xA, self.d_in_fake_A, self.SLOT['D_B'])

    def _adam(self, net, lr, step):
        a = net.arena
        ops.adam(a.p, a.g, a.m, a.v, lr, self.hp['beta1'], 0.999, 1e-8, parallel.grad_scale(self.world_size), step)

    def _allreduce(self, *nets):
        for net in nets:
            parallel.reduce_gradients(net.arena.g, self.world_size)

    # ---- the step ------------------------------------------------------------------------------
    def _part1(self):
        self.losses.zero_()
        self._forward()
        self._phase_G()

    def _part2(self):
        self._adam(self.G_A, self.lr_G, self.step_GA)
        self._adam(self.G_B, self.lr_G, self.step_GB)
        self._pack_generators()

    def _part3(self):
        self._phase_D()

    def _part4(self):
        for D, st in ((self.D_A, self.step_DA), (self.D_B, self.step_DB)):
            self._adam(D, self.lr_D, st)
            D.pack_weights()

    def _query_pools(self):
        """fake_B_pool.query(fake_B) / fake_A_pool.query(fake_A) (cycle_gan_model.py:248-258), in that order.  With a
        pool the discriminator phase always reads the pools' fixed output buffers, so the captured graph stays valid."""
        self.d_in_fake_B = self.pool_B.query(self.GA_real.out)
        self.d_in_fake_A = self.pool_A.query(self.GB_real.out)

    def step(self):
        """optimize_parameters(): four launch segments separated by the gradient all-reduces and the pool query."""
        if self.use_cuda_graph:
            if self._graphs is None:
                self._capture()
            g1, g2, g3, g4 = self._graphs
            g1.replay()
            self._allreduce(self.G_A, self.G_B)
            g2.replay()
            self._query_pools()
            g3.replay()
            self._allreduce(self.D_A, self.D_B)
            g4.replay()
        else:
            self._part1()
            self._allreduce(self.G_A, self.G_B)
            self._part2()
            self._query_pools()
            self._part3()
            self._allreduce(self.D_A, self.D_B)
            self._part4()

    def _mutable_state(self):
        state = []
        for net in (self.G_A, self.G_B, self.D_A, self.D_B):
            state += [t for t in (net.arena.p, net.arena.g, net.arena.m, net.arena.v, net.bufs.p) if t is not None]
        return state + [self.step_GA, self.step_GB, self.step_DA, self.step_DB, self.losses]

    def _tune_pass(self):
        """One eager step on a snapshot of every mutable tensor (each Gemm autotunes on its real operands; the image pools
        are bypassed), then the state is restored so that the captured graphs start from exactly the loaded weights."""
        state = self._mutable_state()
        snap = [t.clone() for t in state]
        gens = self._gens('A') + self._gens('B')
        saved = [g.overlap_wgrad for g in gens]
        for g in gens:
            g.overlap_wgrad = False          # kernels are timed one at a time while tuning
        self._part1()
        self._part2()
        self._part3()
        self._part4()
        torch.cuda.synchronize()
        for g, s in zip(gens, saved):
            g.overlap_wgrad = s
        for t, c in zip(state, snap):
            t.copy_(c)
        self._pack_generators()
        self.D_A.pack_weights()
        self.D_B.pack_weights()
        torch.cuda.synchronize()

    def _capture(self):
        from . import _C
        self._tune_pass()
        if self.pool_A.pool_size > 0:
            self.d_in_fake_B, self.d_in_fake_A = self.pool_B.out, self.pool_A.out
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        graphs = []
        n0 = _C.LAUNCH_COUNT[0]
        with torch.cuda.stream(s):
            for part in (self._part1, self._part2, self._part3, self._part4):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    part()
                graphs.append(g)
        torch.cuda.current_stream().wait_stream(s)
        self._graphs = graphs
        self.launches_per_step = _C.LAUNCH_COUNT[0] - n0

    def get_losses(self):
        """Synchronises, like get_current_losses (models/base_model.py:166-188); values as the reference reports them
        (G terms multiplied by their lambdas, D_A / D_B = (real + fake) * 0.5)."""
        l, hp, S = self.losses.tolist(), self.hp, self.SLOT
        lA, lB, lI = hp['lambda_A'], hp['lambda_B'], hp['lambda_identity']
        return {'D_A': 0.5 * l[S['D_A']], 'G_A': l[S['G_A']], 'G_cycle_A': l[S['G_cycle_A']] * lA,
                'G_idt_A': l[S['G_idt_A']] * lB * lI, 'D_B': 0.5 * l[S['D_B']], 'G_B': l[S['G_B']],
                'G_cycle_B': l[S['G_cycle_B']] * lB, 'G_idt_B': l[S['G_idt_B']] * lA * lI}


class SpadeTrainStep(SpadeDistillStep):
    """backward_G (models/spade_model.py:189-196 -> compute_G_loss, models/modules/spade_modules/spade_model_modules.py:
    97-120) -> optimizer_G.step -> backward_D (-> compute_D_loss :122-139, second no-grad generator forward) ->
    optimizer_D.step: the SPADE distillation step without the frozen teacher and the KA terms.  G_arch['active_fn'] is
    'nn.LeakyReLU' for the reference's training scripts (SPADEModel.modify_commandline_options, spade_model.py:92)."""

    LOSS_NAMES = ['G_gan', 'G_feat', 'G_vgg', 'D_real', 'D_fake']

    def __init__(self, G_arch, D_arch, hp, B, H, W, device='cuda:0', world_size=1, use_cuda_graph=False):
        hp = dict(hp, lambda_distill=0.0)
        super().__init__(None, G_arch, D_arch, hp, B, H, W, device=device, world_size=world_size, use_cuda_graph=use_cuda_graph)
        self.G = self.S

    def load(self, G_sd, D_sd, vgg_sd):
        super().load(None, G_sd, D_sd, vgg_sd)

    def get_losses(self):
        l = super().get_losses()
        return {k: l[k] for k in self.LOSS_NAMES}
