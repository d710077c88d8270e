"""The 'mse' distillation loss (--distill_G_loss_type mse): sum_i F.mse_loss(netA_i(Sact_i), Tact_i) with the 1x1 adaptor
convs netAs (C_S -> C_T, biased) that the reference creates next to the student and trains with optimizer_G
(distillers/base_inception_distiller.py:195-210, inception_distiller.py:111-133; SPADE:
models/modules/spade_modules/spade_distiller_modules.py:17-31, base_spade_distiller_modules.py:70-90).

All on existing kernels: the adaptor is a 1x1 implicit GEMM over the mapped student activation, the loss and its gradient
are catb_recon_loss ('l2' = mean over every element), the adaptor's weight / bias gradients are catb_igemm_wgrad /
catb_channel_sum, and the gradient w.r.t. the student activation is a 1x1 input-gradient GEMM accumulated
(`accumulate` epilogue) into d(activation) at the mapped layer.  Parameters, gradients and Adam moments live in one flat
arena of their own (keys '<i>.weight', '<i>.bias' = netAs[i].state_dict()).
"""
import torch

from . import igemm_plan as P
from . import ops
from .engine import Arena
from .igemm_plan import cpad
from .ops import Act, Gemm


class Adaptors:
    def __init__(self, pairs, device):
        """pairs: [(student Act, C_S, teacher Act, C_T)] in mapping-layer order (real channel counts)."""
        self.dev = device
        self.arena, self.bufs = Arena(with_grad=True), Arena(with_grad=False)
        self.bufs.finalize(device)
        for i, (s, Cs, t, Ct) in enumerate(pairs):
            self.arena.alloc(f'{i}.weight', (Ct, Cs, 1, 1))
            self.arena.alloc(f'{i}.bias', (Ct,))
        self.arena.finalize(device)
        self.layers, self.gemms = [], []
        for i, (s, Cs, t, Ct) in enumerate(pairs):
            assert (s.N, s.H, s.W) == (t.N, t.H, t.W) and s.C == cpad(Cs) and t.C == cpad(Ct)
            B, h, w = s.N, s.H, s.W
            off = self.arena.off(f'{i}.weight')
            L = dict(s=s, t=t, Cs=Cs, Ct=Ct)
            L['out'] = Act.empty(B, h, w, Ct, device, zero=True)         # netA_i(Sact_i)
            L['d_out'] = Act.empty(B, h, w, Ct, device, zero=True)
            L['fwd'] = Gemm(P.Geometry(B, h, w, s.ld, s.coff, h, w, cpad(Ct), 0), P.conv_fprop_units(off, Ct, Cs, 1, 1, 0), Ct, device)
            # d(Sact) += W^T d_out, written with the layout of the generator's d(activation) buffers (full, pitch cpad(C_S))
            L['bwd'] = Gemm(P.Geometry(B, h, w, cpad(Ct), 0, h, w, cpad(Cs), 0), P.conv_dgrad_units(off, Ct, Cs, 1, 1, 0), Cs, device)
            L['bias'], L['dbias'] = self.arena.view(f'{i}.bias'), self.arena.view(f'{i}.bias', 'g')
            self.gemms += [L['fwd'], L['bwd']]
            self.layers.append(L)
        self._packer = None

    def __len__(self):
        return len(self.layers)

    def load_state_dicts(self, sds):
        """sds: [netA.state_dict()] ('weight' [C_T, C_S, 1, 1], 'bias' [C_T])."""
        for i, sd in enumerate(sds):
            self.arena.view(f'{i}.weight').copy_(sd['weight'].to(torch.float32))
            self.arena.view(f'{i}.bias').copy_(sd['bias'].to(torch.float32))
        self.pack_weights()

    def state_dicts(self):
        return [{'weight': self.arena.view(f'{i}.weight').detach().clone().cpu(),
                 'bias': self.arena.view(f'{i}.bias').detach().clone().cpu()} for i in range(len(self.layers))]

    def pack_weights(self):
        if self._packer is None:
            self._packer = ops.PackBatch(self.gemms, self.dev)
        self._packer.run(self.arena.p)

    def loss(self, i, grad_scale, value_slot):
        """value_slot += mse_i (unscaled); d_out = grad_scale * d mse_i / d netA_i(Sact_i); parameter gradients accumulated."""
        L = self.layers[i]
        L['fwd'].fprop(L['s'].t, L['out'].t, bias=L['bias'])
        ops.recon_loss(L['out'], L['t'], L['Ct'], 'l2', grad_scale, value_slot, L['d_out'])
        L['fwd'].wgrad(L['s'].t, L['d_out'].t, self.arena.g)
        ops.channel_sum(L['d_out'], L['dbias'])

    def backward_into(self, i, dact: Act):
        """d(Sact_i) += netA_i^T d_out -- called from the generator's backward pass at mapping layer i."""
        L = self.layers[i]
        assert dact.ld == cpad(L['Cs']) and dact.coff == 0, 'd(activation) buffer with an unexpected layout'
        L['bwd'].fprop(L['d_out'].t, dact.t, accumulate=True)
