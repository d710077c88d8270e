"""Student architecture search of the CAT distillation pipeline: the step just before the hot path.

Restates what `shrink_model` (utils/common.py:315-707) decides -- the channel counts of the pruned student -- without
building ~20 deep copies of the teacher and running profiling forwards through them: the search is a bisection on one
global threshold over the |gamma| of the teacher's normalisation layers (down-sampling norms, the first norm of every
residual branch, up-sampling norms), a channel survives when its |gamma| exceeds the threshold, and a candidate is
scored by the MAC count `model_profiling` would report (utils/model_profiling.py:66-146: convs
cin*cout*k*k*oh*ow/groups, norm layers C*oh*ow unless they track running statistics).  The MAC count is evaluated in
closed form from the architecture dict, so the whole search is host arithmetic on a few thousand floats; the result is
the `arch` dict the engines (cat_b200/engine.py) and the module mirrors (`InceptionGenerator.from_arch`) are compiled from.
The bisection reproduces the reference's fp32 arithmetic (thresholds are fp32 tensors there), so the channel counts are
identical to the reference's (tests/test_prune_cpu.py, fixtures written by the real `shrink`).
"""
import torch


def generator_macs(arch, H, W, parts=False):
    """n_macs of an InceptionGenerator as utils.model_profiling.model_profiling reports it (batch 1); parts=True returns
    (down_sampling, features, up_sampling), the per-section figures trainer.py:121-123 logs."""
    c0, c1, c2, c3, c4 = arch['widths']
    cin, cout, ks = arch['input_nc'], arch['output_nc'], arch['kernel_sizes']
    norm_macs = not arch['track_running_stats']      # model_profiling.py:101-128: norm layers count when they do not track

    def norm(C, h, w):
        return C * h * w if norm_macs else 0
    H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
    m = cin * c0 * 49 * H * W + norm(c0, H, W)
    m += c0 * c1 * 9 * H2 * W2 + norm(c1, H2, W2)
    m += c1 * c2 * 9 * H4 * W4 + norm(c2, H4, W4)
    m_down = m
    hw = H4 * W4
    for blk in arch['blocks']:
        any_branch = False
        for mid, k in zip(blk['res'], ks):
            if mid == 0:
                continue
            any_branch = True
            m += c2 * mid * k * k * hw + norm(mid, H4, W4) + mid * c2 * k * k * hw
        for mid, k in zip(blk['dw'], ks):
            if mid == 0:
                continue
            any_branch = True
            m += c2 * mid * hw + norm(mid, H4, W4) + mid * k * k * hw + norm(mid, H4, W4) + mid * c2 * hw
        # pw_bn exists (and is profiled) for every block; its forward hook only fires when the block has a branch
        if any_branch:
            m += norm(c2, H4, W4)
    m_feat = m - m_down
    m += c2 * c3 * 9 * H2 * W2 + norm(c3, H2, W2)
    m += c3 * c4 * 9 * H * W + norm(c4, H, W)
    m += c4 * cout * 49 * H * W
    if parts:
        return int(m_down), int(m_feat), int(m - m_down - m_feat)
    return int(m)


def _count(gamma, thr):
    return int((gamma.abs() > thr).sum().item())


def _count_lb(gamma, thr, lb):
    """Channels above the threshold, but at least `lb` (the reference then keeps the lb largest, ties included)."""
    n = _count(gamma, thr)
    if n < lb:
        private = torch.sort(gamma.abs().view(-1), descending=True)[0][lb - 1]
        n = int((gamma.abs() >= private).sum().item())
    return n


def _candidate(teacher_sd, arch, thr, lb, ub, ft_lb, final):
    ds = [teacher_sd[f'down_sampling.{i}.weight'] for i in (2, 5, 8)]
    us = [teacher_sd[f'up_sampling.{i}.weight'] for i in (1, 4)]
    widths = []
    for j, g in enumerate(ds):
        if final:
            n = _count_lb(g, thr, lb)
            if j == 0 and n > ub:
                private = torch.sort(g.abs().view(-1), descending=False)[0][int(ub) - 1]
                n = int((g.abs() <= private).sum().item())
            if j == 2 and n < ft_lb:
                n = _count_lb(g, thr, ft_lb)
        else:   # the search loop clamps the counts (utils/common.py:350-359)
            n = max(_count(g, thr), lb)
            if j == 0:
                n = min(n, ub)
            if j == 2:
                n = max(n, ft_lb)
        widths.append(int(n))
    blocks = []
    for i, blk in enumerate(arch['blocks']):
        res, dw = [], []
        jr = jd = 0
        for mid in blk['res']:
            if mid == 0:
                res.append(0)
                continue
            res.append(_count(teacher_sd[f'features.{i}.res_ops.{jr}.1.1.weight'], thr))
            jr += 1
        for mid in blk['dw']:
            if mid == 0:
                dw.append(0)
                continue
            dw.append(_count(teacher_sd[f'features.{i}.dw_ops.{jd}.0.1.weight'], thr))
            jd += 1
        blocks.append({'res': res, 'dw': dw})
    for g in us:
        widths.append(_count_lb(g, thr, lb) if final else max(_count(g, thr), lb))
    return dict(arch, widths=widths, blocks=blocks)


def shrink_arch(teacher_sd, teacher_arch, target_flops, H, W, prune_cin_lb=1, prune_cin_ub=float('inf'), prune_ft_cin_lb=1):
    """The pruned student architecture `shrink_model` would produce for this teacher (state_dict with the norm scales)
    at `target_flops` MACs profiled on an H x W input.  Returns (student_arch, info)."""
    A = teacher_arch
    gammas = [teacher_sd[f'down_sampling.{i}.weight'] for i in (2, 5, 8)]
    for i, blk in enumerate(A['blocks']):      # the first norm of every existing branch (utils/prune.py:19-52)
        gammas += [teacher_sd[f'features.{i}.res_ops.{j}.1.1.weight'] for j in range(sum(1 for c in blk['res'] if c > 0))]
        gammas += [teacher_sd[f'features.{i}.dw_ops.{j}.0.1.weight'] for j in range(sum(1 for c in blk['dw'] if c > 0))]
    gammas += [teacher_sd[f'up_sampling.{i}.weight'] for i in (1, 4)]
    allw = torch.cat([g.detach().float().abs().view(-1) for g in gammas])
    lb, ub = allw.min(), allw.max()              # fp32 scalars, as in the reference
    searched, thr, iters = float('inf'), None, 0
    while bool((ub - lb).abs() > 1e-3 * lb) or searched > target_flops:
        thr = (lb + ub) / 2
        cand = _candidate(teacher_sd, A, thr, prune_cin_lb, prune_cin_ub, prune_ft_cin_lb, final=False)
        searched = generator_macs(cand, H, W)
        if searched > target_flops:
            lb = thr
        else:
            ub = thr
        iters += 1
        if iters > 200:
            raise RuntimeError('shrink_arch: the target (%g MACs) cannot be reached; smallest candidate has %d MACs'
                               % (target_flops, searched))
    student = _candidate(teacher_sd, A, thr, prune_cin_lb, prune_cin_ub, prune_ft_cin_lb, final=True)
    return student, {'threshold': float(thr), 'macs': generator_macs(student, H, W), 'searched_macs': searched, 'iterations': iters}


def shrink(model, opt):
    """Drop-in for `utils.common.shrink(model, opt)` (utils/common.py:872-878) as trainer.py:106-107 calls it."""
    if hasattr(model, 'modules_on_one_gpu'):
        return shrink_spade(model, opt)
    return shrink_inception(model, opt)


def shrink_spade(model, opt):
    """`shrink_spade_model` on a SPADE distiller mirror: `modules_on_one_gpu.netG_student` becomes a freshly initialised
    InceptionSPADEGenerator of the searched architecture, the adaptor convs `netAs` are re-created for the student width
    (utils/common.py:835-843), and the compiled engine is dropped."""
    import copy
    from torch import nn
    from .models import networks
    from .models.spade_networks import InceptionSPADEGenerator
    mm = model.modules_on_one_gpu
    teacher = mm.netG_teacher
    sd = {k: v.detach().float().cpu() for k, v in teacher.state_dict().items() if k.endswith('norm.weight')}
    target = float(getattr(opt, 'target_flops', 0.0))
    assert target > 0, 'opt.target_flops must be positive'
    student_arch, info = shrink_spade_arch(sd, teacher.arch(), target, prune_cin_lb=int(getattr(opt, 'prune_cin_lb', 1)),
                                           prune_cin_ub=getattr(opt, 'prune_cin_ub', float('inf')))
    s_opt = copy.deepcopy(teacher.opt)
    s_opt.ngf = student_arch['fc_out'] // 16
    gpu_ids = list(getattr(model, 'gpu_ids', []))[:1] if getattr(model, 'device', torch.device('cpu')).type == 'cuda' else []
    mm.netG_student = networks.init_net(InceptionSPADEGenerator.from_arch(student_arch, s_opt), opt.init_type, opt.init_gain, gpu_ids)
    mm.netG_student.n_macs = info['macs']
    mm.netG_student.eval()         # shrink_spade_model profiles the new student, which leaves it in eval() (utils/common.py:829)
    teacher.n_macs = spade_generator_macs(teacher.arch())
    ngf_stu = student_arch['fc_out'] // 16
    netAs = nn.ModuleList()
    for layer in mm.mapping_layers:
        fs, ft = (ngf_stu * 16, opt.teacher_ngf * 16) if layer != 'up_1' else (ngf_stu * 4, opt.teacher_ngf * 4)
        netAs.append(nn.Conv2d(fs, ft, kernel_size=1))
    mm.netAs = netAs
    if hasattr(model, 'engine'):
        model.engine = None
        model.__dict__.pop('_engine_cache', None)      # engines compiled for the old architecture
    print('scale threshold: %g, searched flops: %d, target flops: %g' % (info['threshold'], info['macs'], target))
    return info


def shrink_inception(model, opt):
    """`shrink_model` (utils/common.py:315-707) on an Inception distiller mirror: replaces
    `model.netG_student` by a freshly initialised generator of the searched architecture (the reference also copies the
    surviving teacher weights, which trainer.py overwrites with `init_net` on the next line) and drops the compiled
    engine so that the next `set_input` compiles the step for the new shapes.  Returns the search record."""
    from .models import networks
    teacher = model.netG_teacher
    sd = {k: v.detach().float().cpu() for k, v in teacher.state_dict().items() if v.dim() == 1}
    target = float(getattr(opt, 'target_flops', 0.0))
    assert target > 0, 'opt.target_flops must be positive'
    student_arch, info = shrink_arch(sd, teacher.arch(), target, int(opt.data_height), int(opt.data_width),
                                     prune_cin_lb=int(getattr(opt, 'prune_cin_lb', 1)),
                                     prune_cin_ub=getattr(opt, 'prune_cin_ub', float('inf')),
                                     prune_ft_cin_lb=int(getattr(opt, 'prune_ft_cin_lb', 1)))
    gpu_ids = list(getattr(model, 'gpu_ids', []))[:1] if getattr(model, 'device', torch.device('cpu')).type == 'cuda' else []
    model.netG_student = networks.init_net(networks.InceptionGenerator.from_arch(student_arch), opt.init_type, opt.init_gain, gpu_ids)
    model.netG_student.n_macs = info['macs']
    for sec, v in zip(('down_sampling', 'features', 'up_sampling'),
                      generator_macs(student_arch, int(opt.data_height), int(opt.data_width), parts=True)):
        getattr(model.netG_student, sec).n_macs = v        # logged by trainer.py:121-123
    model.netG_student.eval()      # shrink_model profiles the new student, which leaves it in eval() (utils/common.py:148)
    teacher.n_macs = generator_macs(teacher.arch(), int(opt.data_height), int(opt.data_width))
    if getattr(model, 'netAs', None):       # adaptor convs follow the pruned width (utils/common.py:154-161)
        from torch import nn
        dev = model.netAs[0].weight.device
        model.netAs = [nn.Conv2d(student_arch['widths'][2], a.out_channels, kernel_size=1).to(dev) for a in model.netAs]
    if hasattr(model, 'engine'):
        model.engine = None
        model.__dict__.pop('_engine_cache', None)      # engines compiled for the old architecture
    print('scale threshold: %g, searched flops: %d, target flops: %g' % (info['threshold'], info['macs'], target))
    return info


# ----------------------------------------------------------------------------------------------------
# SPADE generator (shrink_spade_model, utils/common.py:710-835)
# ----------------------------------------------------------------------------------------------------
def spade_generator_macs(arch, H=None, W=None):
    """n_macs of an InceptionSPADEGenerator as model_profiling reports it (batch 1): every conv of the six-branch bodies,
    of the SPADE gamma / beta bodies (evaluated on the label map at the block's resolution), the learned shortcuts, fc and
    conv_img; SynchronizedBatchNorm (tracking running statistics), up-sampling and activations count 0."""
    ks, snc = arch['kernel_sizes'], arch['semantic_nc']
    h, w = arch['sh'], arch['sw']
    more = arch['num_upsampling_layers'] in ('more', 'most')
    m = snc * arch['fc_out'] * 9 * h * w

    def body(cin, cout, res, dw, hw):
        t = 0
        for mid, k in zip(res, ks):
            if mid:
                t += cin * mid * k * k * hw + mid * cout * k * k * hw
        for mid, k in zip(dw, ks):
            if mid:
                t += cin * mid * hw + mid * k * k * hw + mid * cout * hw
        return t
    for name in arch['block_names']:
        if name in ('G_middle_0', 'up_0', 'up_1', 'up_2', 'up_3', 'up_4') or (name == 'G_middle_1' and more):
            h, w = 2 * h, 2 * w
        b = arch['blocks'][name]
        hw = h * w
        empty = not any(b['res']) and not any(b['dw'])
        if b['learned_shortcut']:
            m += b['fin'] * b['fout'] * hw
        if not empty:     # forward returns early for a branch-less block: its SPADE body is never run (hooks do not fire)
            m += body(b['fin'], b['fout'], b['res'], b['dw'], hw)
            m += body(snc, 2 * b['fin'], b['spade_res'], b['spade_dw'], hw)
    m += arch['final_nc'] * 3 * 9 * h * w
    return int(m)


def _spade_candidate(sd, A, thr, lb, ub):
    ch_div = 32 if A['num_upsampling_layers'] == 'most' else 16
    c = _count(sd['fc_norm.weight'], thr)
    c = max(c // ch_div, lb) * ch_div
    c = int(min(c // ch_div, ub) * ch_div)
    blocks, cin = {}, c
    for name in A['block_names']:
        T = A['blocks'][name]
        fout = cin // 2 if 'up' in name else cin

        def counts(widths, fmt):
            out, j = [], 0
            for mid in widths:
                if mid == 0:
                    out.append(0)
                    continue
                out.append(_count(sd[fmt.format(name, j)], thr))
                j += 1
            return out
        blocks[name] = {'fin': cin, 'fout': fout, 'res': counts(T['res'], '{}.res_ops.{}.0.norm.weight'),
                        'dw': counts(T['dw'], '{}.dw_ops.{}.0.norm.weight'),
                        'spade_res': counts(T['spade_res'], '{}.spade.res_ops.{}.0.norm.weight'),
                        'spade_dw': counts(T['spade_dw'], '{}.spade.dw_ops.{}.0.norm.weight'), 'learned_shortcut': cin != fout}
        cin = fout
    return dict(A, fc_out=c, final_nc=cin, blocks=blocks)


def shrink_spade_arch(teacher_sd, teacher_arch, target_flops, prune_cin_lb=1, prune_cin_ub=float('inf')):
    """The pruned InceptionSPADEGenerator architecture `shrink_spade_model` would produce (the candidate of the last
    bisection step).  `teacher_arch` carries the latent size (sh, sw) of the profiling resolution."""
    A = teacher_arch
    keys = ['fc_norm.weight']
    for name in A['block_names']:
        T = A['blocks'][name]
        for widths, fmt in ((T['res'], '{}.res_ops.{}.0.norm.weight'), (T['dw'], '{}.dw_ops.{}.0.norm.weight'),
                            (T['spade_res'], '{}.spade.res_ops.{}.0.norm.weight'), (T['spade_dw'], '{}.spade.dw_ops.{}.0.norm.weight')):
            keys += [fmt.format(name, j) for j in range(sum(1 for c in widths if c > 0))]
    allw = torch.cat([teacher_sd[k].detach().float().abs().view(-1) for k in keys])
    lb, ub = allw.min(), allw.max()
    searched, thr, iters, cand = float('inf'), None, 0, None
    while bool((ub - lb).abs() > 1e-3 * lb) or searched > target_flops:
        thr = (lb + ub) / 2
        cand = _spade_candidate(teacher_sd, A, thr, prune_cin_lb, prune_cin_ub)
        searched = spade_generator_macs(cand)
        if searched > target_flops:
            lb = thr
        else:
            ub = thr
        iters += 1
        if iters > 200:
            raise RuntimeError('shrink_spade_arch: the target (%g MACs) cannot be reached; smallest candidate has %d MACs'
                               % (target_flops, searched))
    return cand, {'threshold': float(thr), 'macs': searched, 'iterations': iters}
