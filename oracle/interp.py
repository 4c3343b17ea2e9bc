"""ORACLE (test infrastructure, never shipped): op-by-op fp32 executor for the
three shipped ProgramDesc graphs on torch-CPU.

The reference runs these graphs inside the closed Paddle Inference library
(`predictor_->Run()`, src/ocr_det.cpp:120, src/ocr_cls.cpp:76, src/ocr_rec.cpp:85);
Paddle is not installable here, so this restates the published
operator semantics of Paddle 2.x for the 27 op types the graphs use.
PARITY UNPINNED against a real Paddle run (no Paddle in this environment); the
cls graph runs with the real shipped weights, det/rec with seeded synthetic ones.
"""
from __future__ import annotations
import numpy as np
import torch
import torch.nn.functional as F

from .pdmodel import Program, Op


def _bcast_y(x, y, axis):
    """Paddle elementwise broadcasting: y's dims align to x starting at `axis`
    (axis=-1 → trailing alignment, like numpy)."""
    if x.dim() < y.dim():
        return _bcast_y(y, x, axis)  # paddle swaps when rank(X) < rank(Y)
    if axis == -1 or x.dim() == y.dim():
        return y
    shape = [1] * x.dim()
    for i, d in enumerate(y.shape):
        shape[axis + i] = d
    return y.reshape(shape)


def _pool_out(size, k, pad, stride, ceil_mode):
    # paddle/phi/kernels/funcs/pooling.h PoolOutputSize: C++ int division (truncates toward zero)
    num = size - k + 2 * pad + (stride - 1 if ceil_mode else 0)
    return int(num / stride) + 1


def _pool2d(x, a):
    ptype = a['pooling_type']
    k = list(a['ksize']); s = list(a['strides']); p = list(a['paddings'])
    if a.get('global_pooling', False) or (a.get('adaptive', False) and k == [1, 1]):
        return x.mean(dim=(2, 3), keepdim=True) if ptype == 'avg' else x.amax(dim=(2, 3), keepdim=True)
    assert not a.get('adaptive', False)
    assert len(p) == 2
    N, C, H, W = x.shape
    oh = _pool_out(H, k[0], p[0], s[0], a.get('ceil_mode', False))
    ow = _pool_out(W, k[1], p[1], s[1], a.get('ceil_mode', False))
    out = x.new_empty((N, C, oh, ow))
    excl = a.get('exclusive', True)
    for i in range(oh):
        h0 = i * s[0] - p[0]; h1 = min(h0 + k[0], H + p[0])
        hs, he = max(h0, 0), min(h1, H)
        for j in range(ow):
            w0 = j * s[1] - p[1]; w1 = min(w0 + k[1], W + p[1])
            ws, we = max(w0, 0), min(w1, W)
            win = x[:, :, hs:he, ws:we]
            if ptype == 'max':
                out[:, :, i, j] = win.amax(dim=(2, 3))
            else:
                cnt = (he - hs) * (we - ws) if excl else (h1 - h0) * (w1 - w0)
                out[:, :, i, j] = win.sum(dim=(2, 3)) / cnt
    return out


def _reshape_target(x, shape):
    shape = list(shape)
    for i, d in enumerate(shape):
        if d == 0:
            shape[i] = x.shape[i]
    return shape


def run_program(prog: Program, params: dict, feed, keep=None, want_all=False, grad=False):
    """Execute `prog` on `feed` (NCHW fp32). Returns (output ndarray, {name: ndarray} for `keep`).
    grad=True (tools/train_synth_det.py only): `params` / `feed` are torch tensors, autograd stays on and the
    output is returned as a torch tensor."""
    if grad:
        env = dict(params)
    else:
        env = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()}
    kept = {}
    out = None
    with (torch.enable_grad() if grad else torch.no_grad()):
        for op in prog.ops:
            t, a = op.type, op.attrs
            if t == 'feed':
                env[op.o('Out')] = feed if grad else torch.from_numpy(np.ascontiguousarray(feed, dtype=np.float32)); continue
            if t == 'fetch':
                out = env[op.i('X')]; continue
            if t in ('conv2d', 'depthwise_conv2d'):
                assert a.get('padding_algorithm', 'EXPLICIT') == 'EXPLICIT' and a.get('data_format', 'NCHW') == 'NCHW'
                p = a['paddings']; assert len(p) == 2
                r = F.conv2d(env[op.i('Input')], env[op.i('Filter')], None, tuple(a['strides']),
                             tuple(p), tuple(a['dilations']), a['groups'])
                env[op.o('Output')] = r
            elif t == 'conv2d_transpose':
                p = a['paddings']; assert len(p) == 2 and not a.get('output_size') and not a.get('output_padding')
                r = F.conv_transpose2d(env[op.i('Input')], env[op.i('Filter')], None, tuple(a['strides']),
                                       tuple(p), 0, a['groups'], tuple(a['dilations']))
                env[op.o('Output')] = r
            elif t == 'batch_norm':
                x = env[op.i('X')]
                sc, b, m, v = (env[op.i(k)] for k in ('Scale', 'Bias', 'Mean', 'Variance'))
                sh = [1, -1] + [1] * (x.dim() - 2)
                inv = 1.0 / torch.sqrt(v + a['epsilon'])
                env[op.o('Y')] = (x - m.reshape(sh)) * (inv * sc).reshape(sh) + b.reshape(sh)
            elif t in ('elementwise_add', 'elementwise_mul'):
                x, y = env[op.i('X')], env[op.i('Y')]
                if x.dim() >= y.dim():
                    yy = _bcast_y(x, y, a.get('axis', -1)); xx = x
                else:
                    xx = _bcast_y(y, x, a.get('axis', -1)); yy = y
                env[op.o('Out')] = xx + yy if t == 'elementwise_add' else xx * yy
            elif t == 'hard_swish':
                x = env[op.i('X')]
                env[op.o('Out')] = x * torch.clamp(x + a['offset'], 0.0, a['threshold']) / a['scale']
            elif t == 'hard_sigmoid':
                env[op.o('Out')] = torch.clamp(env[op.i('X')] * a['slope'] + a['offset'], 0.0, 1.0)
            elif t == 'relu':
                env[op.o('Out')] = torch.relu(env[op.i('X')])
            elif t == 'sigmoid':
                env[op.o('Out')] = torch.sigmoid(env[op.i('X')])
            elif t == 'swish':
                x = env[op.i('X')]
                env[op.o('Out')] = x * torch.sigmoid(a.get('beta', 1.0) * x)
            elif t == 'pool2d':
                env[op.o('Out')] = _pool2d(env[op.i('X')], a)
            elif t == 'nearest_interp_v2':
                x = env[op.i('X')]
                assert a['interp_method'] == 'nearest' and not a['align_corners']
                sc = a['scale']; oh, ow = int(x.shape[2] * sc[0]), int(x.shape[3] * sc[1])
                iy = torch.floor(torch.arange(oh, dtype=torch.float32) * (x.shape[2] / oh)).long()
                ix = torch.floor(torch.arange(ow, dtype=torch.float32) * (x.shape[3] / ow)).long()
                env[op.o('Out')] = x[:, :, iy][:, :, :, ix]
            elif t == 'concat':
                env[op.o('Out')] = torch.cat([env[n] for n in op.inputs['X']], dim=a['axis'])
            elif t == 'shape':
                env[op.o('Out')] = torch.tensor(list(env[op.i('Input')].shape), dtype=torch.int64)
            elif t == 'slice':
                x = env[op.i('Input')]
                idx = [slice(None)] * x.dim()
                for ax, st, en in zip(a['axes'], a['starts'], a['ends']):
                    idx[ax] = slice(st, min(en, x.shape[ax]))
                r = x[tuple(idx)]
                for ax in sorted(a.get('decrease_axis', []), reverse=True):
                    r = r.squeeze(ax) if r.dim() > 1 else r
                env[op.o('Out')] = r
            elif t == 'fill_constant':
                env[op.o('Out')] = torch.full(a['shape'], a['value'])
            elif t == 'reshape2':
                x = env[op.i('X')]
                env[op.o('Out')] = x.reshape(_reshape_target(x, a['shape']))
            elif t == 'matmul_v2':
                x, y = env[op.i('X')], env[op.i('Y')]
                if a.get('trans_x'): x = x.transpose(-1, -2)
                if a.get('trans_y'): y = y.transpose(-1, -2)
                env[op.o('Out')] = torch.matmul(x, y)
            elif t == 'softmax':
                env[op.o('Out')] = torch.softmax(env[op.i('X')], dim=a['axis'])
            elif t in ('assign', 'dropout'):
                if t == 'dropout':
                    assert a.get('is_test', True) and a.get('dropout_implementation') == 'upscale_in_train', a
                env[op.o('Out')] = env[op.i('X')]
            elif t == 'flatten_contiguous_range':
                env[op.o('Out')] = torch.flatten(env[op.i('X')], a['start_axis'], a['stop_axis'])
            elif t == 'transpose2':
                env[op.o('Out')] = env[op.i('X')].permute(*a['axis']).contiguous()
            elif t == 'layer_norm':
                x = env[op.i('X')]; bna = a['begin_norm_axis']
                env[op.o('Y')] = F.layer_norm(x, tuple(x.shape[bna:]), env[op.i('Scale')].reshape(x.shape[bna:]),
                                              env[op.i('Bias')].reshape(x.shape[bna:]), a['epsilon'])
            elif t == 'scale':
                x = env[op.i('X')]
                env[op.o('Out')] = (x * a['scale'] + a['bias']) if a.get('bias_after_scale', True) else (x + a['bias']) * a['scale']
            elif t == 'squeeze2':
                x = env[op.i('X')]
                for ax in sorted(a['axes'], reverse=True):
                    x = x.squeeze(ax)
                env[op.o('Out')] = x
            else:
                raise NotImplementedError(t)
            if keep or want_all:
                for k, names in op.outputs.items():
                    for n in names:
                        if n in env and (want_all or n in keep) and k in ('Out', 'Output', 'Y'):
                            kept[n] = env[n].detach().numpy()
    return (out if grad else out.numpy()), kept
