"""CPU oracle: functional restatement of the reference's Dense-U-Net hot path.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; the product path
(dmmfods_b200/) never does and fails loudly when its CUDA library is missing.

Parity status: the reference (p-mc-grath/DMMFODS) ships NO tests / golden vectors
(SURVEY.md section 4), so nothing of its own pins results.  This oracle is pinned instead
against outputs of the unmodified reference itself, executed in the authoring container via
oracle/ref_shim.py: tests/golden/make_golden.py writes tests/golden/*.npz and
tests/test_oracle_golden.py checks this file against them (and, when /root/reference is
present, against the live reference).

Third-party arithmetic: the dense-block math lives in torchvision (un-pinned dependency,
requirements.txt:8; reference imports `_DenseLayer/_DenseBlock/_Transition` at
Dense_U_Net_lidar.py:9).  It is restated below from the published algorithm
(torchvision 0.26 models/densenet.py:31-133, "tv:" citations), not imported.

Everything is a pure function of (state_dict, config, inputs) written with
torch.nn.functional CPU ops in the dtype of the inputs (float32 = reference arithmetic,
float64 = high-precision yard-stick).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default (reference never overrides it)
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default


class _RoundBf16(torch.autograd.Function):
    """Storage-precision emulation of the CUDA path (NOT reference arithmetic): the value is rounded to
    bfloat16 in forward and the gradient flowing back through the same point is rounded as well - the CUDA
    engine stores exactly these tensors (conv outputs, BN-ReLU outputs and their gradients) as bf16."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.to(torch.bfloat16).to(x.dtype) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(g.dtype) if ctx.bwd else g), None, None


class _Emu:
    """emulate=False: plain reference arithmetic.  emulate=True: bf16 rounding at the CUDA engine's storage points."""

    def __init__(self, emulate):
        self.on = bool(emulate)

    def act(self, x):            # stored activation (and its stored gradient)
        return _RoundBf16.apply(x, True, True) if self.on else x

    def val(self, x):            # operand rounded when read (weights, conv0 input): gradient stays fp32
        return _RoundBf16.apply(x, True, False) if self.on else x

    def grad(self, x):           # fp32 forward value whose incoming gradient is stored as bf16 (logits)
        return _RoundBf16.apply(x, False, True) if self.on else x


_EMU = _Emu(False)


def fusion_mode(model_cfg):
    """Dense_U_Net_lidar.py:57-65."""
    cb = model_cfg["concat_before_block_num"]
    c2 = model_cfg["stream_2_in_channels"]
    if cb == 1 and c2 == 0:
        return "no"
    if cb == 1 and c2 > 0:
        return "early"
    if 1 < cb <= len(model_cfg["block_config"]):
        return "mid"
    raise AttributeError("invalid fusion configuration")


def _bn_relu(x, sd, prefix, train, new_stats):
    """nn.BatchNorm2d + nn.ReLU.  Training: batch mean / biased var normalise, running stats
    updated with momentum 0.1 and the UNBIASED variance, num_batches_tracked += 1."""
    rm = sd[prefix + ".running_mean"].to(x.dtype).clone()
    rv = sd[prefix + ".running_var"].to(x.dtype).clone()
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"].to(x.dtype), sd[prefix + ".bias"].to(x.dtype),
                     training=train, momentum=BN_MOMENTUM, eps=BN_EPS)
    if train and new_stats is not None:
        new_stats[prefix + ".running_mean"] = rm
        new_stats[prefix + ".running_var"] = rv
        new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    return F.relu(y)


def _bn_relu_r(x, sd, prefix, train, new_stats):
    return _EMU.act(_bn_relu(x, sd, prefix, train, new_stats))


def _conv(x, sd, name, stride=1, padding=0, store=True):
    y = F.conv2d(x, _EMU.val(sd[name + ".weight"].to(x.dtype)), None, stride=stride, padding=padding)
    return _EMU.act(y) if store else _EMU.grad(y)


def _dense_layer(feats, sd, p, train, ns):
    """tv:_DenseLayer.forward (tv:76-93) with bn_function (tv:47-50); drop_rate == 0."""
    x = torch.cat(feats, 1)
    b = _conv(_bn_relu_r(x, sd, p + ".norm1", train, ns), sd, p + ".conv1")           # 1x1 -> bn_size*k
    return _conv(_bn_relu_r(b, sd, p + ".norm2", train, ns), sd, p + ".conv2", padding=1)  # 3x3 -> k


def _dense_block(x, sd, p, num_layers, train, ns):
    """tv:_DenseBlock.forward (tv:119-124)."""
    feats = [x]
    for i in range(num_layers):
        feats.append(_dense_layer(feats, sd, "%s.denselayer%d" % (p, i + 1), train, ns))
    return torch.cat(feats, 1)


def _transition(x, sd, p, train, ns):
    """tv:_Transition (tv:127-133): BN -> ReLU -> 1x1 conv -> AvgPool2d(2, 2)."""
    a = _bn_relu(x, sd, p + ".norm", train, ns)
    if _EMU.on:      # the CUDA path pools first (pool and 1x1 conv commute) and stores the pooled activation
        return _conv(_EMU.act(F.avg_pool2d(a, 2, 2)), sd, p + ".conv")
    return F.avg_pool2d(_conv(a, sd, p + ".conv"), 2, 2)


def _stem(x, sd, p, train, ns, trace=None):
    """features.conv0/norm0/relu0/pool0 (Dense_U_Net_lidar.py:72-78 and 156-162)."""
    x = _conv(_EMU.val(x), sd, p + ".conv0", stride=2, padding=3)
    x = _bn_relu(x, sd, p + ".norm0", train, ns)
    size_after_relu0 = x.shape
    x = _EMU.val(F.max_pool2d(x, kernel_size=3, stride=2, padding=1))   # its gradient stays fp32 in the engine
    return x, size_after_relu0


def oracle_forward(sd, model_cfg, stream_1_data, stream_2_data, train=True, trace=None, emulate_bf16=False):
    """Forward pass of Dense_U_Net_lidar (Dense_U_Net_lidar.py:210-267).

    sd: state_dict (torch tensors); model_cfg: dict with the keys of helper:111-123.
    Returns (logits, new_stats) - new_stats holds the updated BN buffers in training mode.
    `trace`, if a dict, receives named intermediate activations for per-stage parity tests.
    """
    global _EMU
    _EMU = _Emu(emulate_bf16)
    fusion = fusion_mode(model_cfg)
    block_config = tuple(model_cfg["block_config"])
    cb = model_cfg["concat_before_block_num"]
    ns = OrderedDict()
    nb = len(block_config)

    def rec(name, t):
        if trace is not None:
            trace[name] = t

    # Dense_U_Net_lidar.py:224-235
    if fusion == "no":
        skip0 = stream_1_data
        x = stream_1_data
    else:
        skip0 = torch.cat((stream_1_data, stream_2_data), 1)
        x = skip0 if fusion == "early" else stream_1_data

    s2 = None
    if fusion == "mid":                                   # :233 stream_2_features(lidar)
        s2, _ = _stem(stream_2_data, sd, "stream_2_features", train, ns)
        for b in range(cb - 1):
            s2 = _dense_block(s2, sd, "stream_2_features.denseblock%d" % (b + 1), block_config[b], train, ns)
            s2 = _transition(s2, sd, "stream_2_features.transition%d" % (b + 1), train, ns)
        rec("stream_2", s2)

    # encoder :238-252
    x, size0 = _stem(x, sd, "features", train, ns)
    rec("stem", x)
    sizes = [size0]
    skips = [skip0]
    for b in range(nb):
        x = _dense_block(x, sd, "features.denseblock%d" % (b + 1), block_config[b], train, ns)
        rec("block%d" % (b + 1), x)
        if b != nb - 1:
            skips.append(x)
            sizes.append(x.shape)
            x = _transition(x, sd, "features.transition%d" % (b + 1), train, ns)
            rec("transition%d" % (b + 1), x)
            if fusion == "mid" and b + 1 == cb - 1:       # concat after transition cb-1 (:53, :242-245)
                assert x.shape == s2.shape, "%s %s" % (x.shape, s2.shape)
                x = torch.cat((x, s2), 1)
                x = _conv(_bn_relu_r(x, sd, "concat_module.norm", train, ns), sd, "concat_module.conv")
                rec("concat_module", x)

    # decoder :255-261
    for k in range(1, nb + 1):
        p = "decoder.Transposed_Convolution_Sequence_%d" % k
        if k > 1:
            x = torch.cat((x, skips.pop()), 1)
        x = _bn_relu_r(x, sd, p + ".norm0", train, ns)
        x = _conv(x, sd, p + ".conv_reduce")
        x = _bn_relu_r(x, sd, p + ".norm1", train, ns)
        tgt = sizes.pop()
        w = _EMU.val(sd["decoder.Transposed_Convolution_%d.weight" % k].to(x.dtype))
        # nn.ConvTranspose2d(C, C, 3, stride=2, padding=1)(x, output_size=tgt): output_padding derived
        oph = tgt[2] - ((x.shape[2] - 1) * 2 - 2 + 3)
        opw = tgt[3] - ((x.shape[3] - 1) * 2 - 2 + 3)
        x = _EMU.act(F.conv_transpose2d(x, w, None, stride=2, padding=1, output_padding=(oph, opw)))
        rec("dec%d" % k, x)
    x = F.interpolate(x, scale_factor=2, mode="nearest")    # nn.Upsample(scale_factor=2) :120

    # head :264-265
    x = torch.cat((x, skips.pop()), 1)
    p = "dec_out_to_heat_maps"
    x = _conv(_bn_relu_r(x, sd, p + ".norm0", train, ns), sd, p + ".refine0", padding=1)
    rec("refine0", x)
    x = _conv(_bn_relu_r(x, sd, p + ".norm1", train, ns), sd, p + ".refine1", padding=2, store=False)
    _EMU = _Emu(False)
    return x, ns


def bce_with_logits(x, t):
    """torch.nn.BCEWithLogitsLoss(reduction='none') (Agent.py:54,247):
    l = max(x,0) - x*t + log1p(exp(-|x|)); dl/dx = sigmoid(x) - t."""
    return torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-torch.abs(x)))


def oracle_train_step(sd, model_cfg, stream_1_data, stream_2_data, target, dtype=torch.float32, emulate_bf16=False):
    """forward + BCE(reduction none) + backward(ones) (Agent.py:244-264).
    Returns dict(logits, loss, loss_per_class, grads{name: tensor}, new_stats)."""
    params = OrderedDict()
    full = OrderedDict()
    for k, v in sd.items():
        if v.is_floating_point():
            v = v.detach().to(dtype)
            if not (k.endswith("running_mean") or k.endswith("running_var")):
                v = v.clone().requires_grad_(True)
                params[k] = v
        full[k] = v
    logits, ns = oracle_forward(full, model_cfg, stream_1_data.to(dtype), stream_2_data.to(dtype), train=True,
                                emulate_bf16=emulate_bf16)
    loss = bce_with_logits(logits, target.to(dtype))
    loss.backward(torch.ones_like(loss))
    return {
        "logits": logits.detach(),
        "loss": loss.detach(),
        "loss_per_class": loss.detach().sum(dim=(0, 2, 3)),
        "grads": OrderedDict((k, p.grad) for k, p in params.items()),
        "new_stats": ns,
    }
