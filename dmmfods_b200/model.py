"""Drop-in replacement of dmmfods/graphs/models/Dense_U_Net_lidar.py (the reference's L2 model layer).

Same constructor (`Dense_U_Net_lidar(config)`), factories (`densenet{121,161,169,201}_u_lidar(pretrained,
progress, config)`), `forward(stream_1_data, stream_2_data) -> logits`, public attributes and - key for the
agents/ training loop and for torchvision checkpoints - the same module tree, hence the same `state_dict()`
keys, shapes and dtypes and the same parameter initialisation sequence (identical
values under the same torch seed).  The torch.nn modules below only HOLD parameters; all arithmetic runs in
the hand-written sm_100a kernels through `engine.Engine`.  There is no CPU / PyTorch fallback: calling the
model with CPU tensors or without the CUDA library raises.
"""
import os
from collections import OrderedDict, deque

import torch
import torch.nn as nn

from .config import get_config
from .engine import Engine


class _DenseLayer(nn.Module):
    """parameter container with torchvision's `_DenseLayer` sub-module names (tv:31-45)."""

    def __init__(self, num_input_features, growth_rate, bn_size, drop_rate, memory_efficient=False):
        super().__init__()
        self.norm1 = nn.BatchNorm2d(num_input_features)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(num_input_features, bn_size * growth_rate, kernel_size=1, stride=1, bias=False)
        self.norm2 = nn.BatchNorm2d(bn_size * growth_rate)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(bn_size * growth_rate, growth_rate, kernel_size=3, stride=1, padding=1, bias=False)
        self.drop_rate = float(drop_rate)
        self.memory_efficient = memory_efficient


class _DenseBlock(nn.ModuleDict):
    """`denselayer%d` children like torchvision's `_DenseBlock` (tv:96-117)."""

    def __init__(self, num_layers, num_input_features, bn_size, growth_rate, drop_rate, memory_efficient=False):
        super().__init__()
        for i in range(num_layers):
            self.add_module("denselayer%d" % (i + 1),
                            _DenseLayer(num_input_features + i * growth_rate, growth_rate, bn_size, drop_rate,
                                        memory_efficient))


class _Transition(nn.Sequential):
    """norm / relu / conv / pool like torchvision's `_Transition` (tv:127-133)."""

    def __init__(self, num_input_features, num_output_features):
        super().__init__()
        self.norm = nn.BatchNorm2d(num_input_features)
        self.relu = nn.ReLU(inplace=True)
        self.conv = nn.Conv2d(num_input_features, num_output_features, kernel_size=1, stride=1, bias=False)
        self.pool = nn.AvgPool2d(kernel_size=2, stride=2)


class _NetFn(torch.autograd.Function):
    """autograd bridge: forward = engine forward program, backward = engine backward program.

    The parameter gradients handed to autograd are VIEWS of the engine's flat fp32 gradient buffer (no per-tensor copies: 510
    tensors on DenseNet-121).  autograd's AccumulateGrad adopts such a view as `p.grad` when `p.grad` is None (the state
    after `optimizer.zero_grad()`), which is what the reference loop does every step (Agent.py:263).  If gradients are being
    ACCUMULATED over several backward passes (`p.grad` already set), aliased `p.grad`s are detached from the buffer first and
    copies are returned, so `p.grad += g` keeps its meaning."""

    @staticmethod
    def forward(ctx, model, eng, x1, x2, *params):
        ctx.eng = eng
        ctx.model = model
        out = eng.forward(x1, x2)
        ctx.version = eng.fwd_count
        # the logits buffer is engine-owned and rewritten by the next forward: hand out a copy (236 MB at 32x3x640x960:
        # ~0.07 ms of an ~80 ms step) so that a caller may keep predictions of earlier steps, as with the reference module
        return out.clone()

    @staticmethod
    def backward(ctx, dlogits):
        eng = ctx.eng
        if ctx.version != eng.fwd_count:
            raise RuntimeError("dmmfods_b200: backward() called after another forward() of the same shape "
                               "(activations are kept in static buffers: one outstanding graph per shape)")
        plist = ctx.model._param_list
        lo = eng.gflat.data_ptr()
        hi = lo + eng.gflat.numel() * 4
        accumulate = False
        for p in plist:
            g = p.grad
            if g is not None:
                accumulate = True
                if lo <= g.data_ptr() < hi:
                    p.grad = g.clone()          # un-alias before the engine rewrites its buffer
        grads = eng.backward(dlogits.contiguous())
        names = ctx.model._param_order
        if accumulate:
            outs = tuple(grads[n].clone() for n in names)
        else:
            outs = tuple(grads[n].view(grads[n].shape) for n in names)     # fresh view objects: adoptable by AccumulateGrad
        return (None, None, None, None) + outs


class Dense_U_Net_lidar(nn.Module):
    """U-Net-like heat-map network with a DenseNet encoder and an optional LiDAR stream
    (Dense_U_Net_lidar.py:18-267); see the module docstring."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        m = config.model
        self.growth_rate = m.growth_rate
        self.block_config = m.block_config
        self.num_init_features = m.num_init_features
        self.bn_size = m.bn_size
        self.drop_rate = m.drop_rate
        self.memory_efficient = m.memory_efficient
        self.num_classes = m.num_classes
        self.concat_before_block_num = m.concat_before_block_num
        self.num_layers_before_blocks = m.num_layers_before_blocks
        self.concat_after_module_idx = self.num_layers_before_blocks - 1 + 2 * (self.concat_before_block_num - 1)
        self.stream_1_in_channels = m.stream_1_in_channels
        self.stream_2_in_channels = m.stream_2_in_channels
        self.network_input_channels = self.stream_1_in_channels
        if self.concat_before_block_num == 1 and self.stream_2_in_channels == 0:
            self.fusion = "no"
        elif self.concat_before_block_num == 1 and self.stream_2_in_channels > 0:
            self.fusion = "early"
            self.network_input_channels += self.stream_2_in_channels
        elif 1 < self.concat_before_block_num <= len(self.block_config):
            self.fusion = "mid"
        else:
            raise AttributeError

        def stem(cin):
            return nn.Sequential(OrderedDict([
                ("conv0", nn.Conv2d(cin, self.num_init_features, kernel_size=7, stride=2, padding=3, bias=False)),
                ("norm0", nn.BatchNorm2d(self.num_init_features)),
                ("relu0", nn.ReLU(inplace=True)),
                ("pool0", nn.MaxPool2d(kernel_size=3, stride=2, padding=1)),
            ]))

        def dense(num_layers, num_features):
            return _DenseBlock(num_layers=num_layers, num_input_features=num_features, bn_size=self.bn_size,
                               growth_rate=self.growth_rate, drop_rate=self.drop_rate,
                               memory_efficient=self.memory_efficient)

        # encoder (stream_1): DenseNet without norm5 / classifier
        self.features = stem(self.network_input_channels)
        sizes = deque([self.num_init_features + 2 * self.growth_rate])
        nf = self.num_init_features
        for i, num_layers in enumerate(self.block_config):
            self.features.add_module("denseblock%d" % (i + 1), dense(num_layers, nf))
            nf += num_layers * self.growth_rate
            sizes.append(nf)
            if i != len(self.block_config) - 1:
                self.features.add_module("transition%d" % (i + 1), _Transition(nf, nf // 2))
                nf //= 2

        # decoder: 1x1 reduce + stride-2 transposed conv per level, skip concatenation in forward
        self.decoder = nn.Sequential()
        num_in = sizes.pop()
        for i in range(len(self.block_config)):
            nf = sizes.pop()
            self.decoder.add_module("Transposed_Convolution_Sequence_%d" % (i + 1), nn.Sequential(OrderedDict([
                ("norm0", nn.BatchNorm2d(num_in)),
                ("relu0", nn.ReLU(inplace=True)),
                ("conv_reduce", nn.Conv2d(num_in, nf, kernel_size=1, stride=1, padding=0, bias=False)),
                ("norm1", nn.BatchNorm2d(nf)),
                ("relu1", nn.ReLU(inplace=True)),
            ])))
            self.decoder.add_module("Transposed_Convolution_%d" % (i + 1),
                                    nn.ConvTranspose2d(nf, nf, 3, stride=2, padding=1, bias=False))
            num_in = nf * 2
        self.decoder.add_module("Upsampling", nn.Upsample(scale_factor=2))

        cin_head = nf + self.stream_1_in_channels + self.stream_2_in_channels
        self.dec_out_to_heat_maps = nn.Sequential(OrderedDict([
            ("norm0", nn.BatchNorm2d(cin_head)),
            ("relu0", nn.ReLU(inplace=True)),
            ("refine0", nn.Conv2d(cin_head, nf // 2, 3, stride=1, padding=1, bias=False)),
            ("norm1", nn.BatchNorm2d(nf // 2)),
            ("relu1", nn.ReLU(inplace=True)),
            ("refine1", nn.Conv2d(nf // 2, self.num_classes, 5, stride=1, padding=2, bias=False)),
        ]))

        if self.fusion == "mid":
            self.stream_2_features = stem(self.stream_2_in_channels)
            nf = self.num_init_features
            for i, num_layers in enumerate(self.block_config):
                if i == self.concat_before_block_num - 1:
                    break
                self.stream_2_features.add_module("denseblock%d" % (i + 1), dense(num_layers, nf))
                nf += num_layers * self.growth_rate
                if i != len(self.block_config) - 1:
                    self.stream_2_features.add_module("transition%d" % (i + 1), _Transition(nf, nf // 2))
                    nf //= 2
            nf = self.features[self.concat_after_module_idx + 1].denselayer1.norm1.num_features
            self.concat_module = nn.Sequential(OrderedDict([
                ("norm", nn.BatchNorm2d(nf * 2)),
                ("relu", nn.ReLU(inplace=True)),
                ("conv", nn.Conv2d(nf * 2, nf, kernel_size=1, stride=1, padding=0, bias=False)),
            ]))

        for mod in self.modules():
            if isinstance(mod, nn.Conv2d):
                nn.init.kaiming_normal_(mod.weight)
            elif isinstance(mod, nn.BatchNorm2d):
                nn.init.constant_(mod.weight, 1)
                nn.init.constant_(mod.bias, 0)
            elif isinstance(mod, nn.Linear):
                nn.init.constant_(mod.bias, 0)

        self.num_params = sum(p.numel() for p in self.parameters())
        self._engines = {}
        self._param_order = None
        self._param_list = None

    # ------------------------------------------------------------------------------------------------
    def model_cfg(self):
        return {"growth_rate": self.growth_rate, "block_config": tuple(self.block_config),
                "num_init_features": self.num_init_features, "bn_size": self.bn_size, "drop_rate": self.drop_rate,
                "num_classes": self.num_classes, "concat_before_block_num": self.concat_before_block_num,
                "stream_1_in_channels": self.stream_1_in_channels, "stream_2_in_channels": self.stream_2_in_channels}

    def engine(self, B, H, W, training=None, need_backward=True, precision="bf16", bucket_bytes=32 << 20):
        """the (cached) execution engine for one input shape; rebuilt when parameters were re-allocated.
        precision="tf32": the strict forward mode (fp32 storage, kind::tf32 MMAs; forward + loss only)."""
        training = self.training if training is None else training
        if precision in ("tf32", "tf32x3"):
            need_backward = False
        sd = self.state_dict(keep_vars=True)
        dev = next(iter(sd.values())).device
        if dev.type != "cuda":
            raise RuntimeError("dmmfods_b200: the model must live on a CUDA (sm_100) device - there is no CPU path")
        sig = tuple(v.data_ptr() for v in sd.values())
        key = (B, H, W, bool(training), bool(need_backward and training), precision, int(bucket_bytes))
        ent = self._engines.get(key)
        if ent is None or ent[1] != sig:
            for v in sd.values():
                if v.is_floating_point() and (v.dtype != torch.float32 or not v.is_contiguous()):
                    raise TypeError("dmmfods_b200: parameters and BN buffers must be contiguous float32")
            params = OrderedDict((k, v.data if isinstance(v, nn.Parameter) else v) for k, v in sd.items())
            # drop engines of stale parameter storage (their buffers would otherwise stay alive)
            self._engines = {k: e for k, e in self._engines.items() if e[1] == sig}
            eng = Engine(params, self.model_cfg(), B, H, W, training=training, need_backward=need_backward, precision=precision,
                         bucket_bytes=bucket_bytes)
            eng.fwd_count = 0
            self._engines[key] = (eng, sig)
            ent = self._engines[key]
        self._param_order = [k for k, _ in self.named_parameters()]
        self._param_list = [p for _, p in self.named_parameters()]
        return ent[0]

    def forward(self, stream_1_data, stream_2_data):
        """stream_1_data: (B, stream_1_in_channels, H, W) RGB; stream_2_data: (B, stream_2_in_channels, H, W)
        LiDAR (ignored for fusion 'no').  Returns raw logits (B, num_classes, H, W) (:210-267)."""
        if not stream_1_data.is_cuda:
            raise RuntimeError("dmmfods_b200: inputs must be CUDA tensors - there is no CPU path")
        B, c1, H, W = stream_1_data.shape
        if c1 != self.stream_1_in_channels:
            raise RuntimeError("expected %d stream_1 channels, got %d" % (self.stream_1_in_channels, c1))
        if self.fusion != "no":
            assert stream_2_data.shape == (B, self.stream_2_in_channels, H, W), \
                "stream_2_data %s does not match stream_1_data %s" % (tuple(stream_2_data.shape),
                                                                     tuple(stream_1_data.shape))
        x1 = stream_1_data.detach().to(torch.float32)
        x2 = stream_2_data.detach().to(torch.float32) if self.fusion != "no" else None
        grad_on = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters())
        eng = self.engine(B, H, W, need_backward=grad_on)
        eng.fwd_count += 1
        if not grad_on:
            return eng.forward(x1, x2).clone()
        return _NetFn.apply(self, eng, x1, x2, *self.parameters())

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_engines"] = {}
        d["_param_list"] = None
        return d

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, {} if k == "_engines" else (None if k == "_param_list" else copy.deepcopy(v, memo)))
        return new


class FusedBCEWithLogits(nn.Module):
    """torch.nn.BCEWithLogitsLoss(reduction='none') (Agent.py:54) on the fused CUDA kernel: returns the
    un-reduced (B,C,H,W) loss; its backward for any cotangent g is g * (sigmoid(x) - t)."""

    def forward(self, prediction, target):
        return _BCEFn.apply(prediction, target)


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t):
        from . import ops
        x = x.contiguous()
        t = t.to(torch.float32).contiguous()
        loss = torch.empty_like(x)
        grad = torch.empty_like(x)
        ops.bce_logits(x, t, loss, grad, None)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return g * grad, None


def _load_state_dict(model, config, model_url, progress):
    """torchvision ImageNet weights -> stream_1 (and stream_2 except conv0) (Dense_U_Net_lidar.py:269-309).
    Needs network access for `load_state_dict_from_url`, exactly like the reference."""
    import re
    pattern = re.compile(r"^(.*denselayer\d+\.(?:norm|relu|conv))\.((?:[12])\.(?:weight|bias|running_mean|running_var))$")
    sd_tv = torch.hub.load_state_dict_from_url(model_url, progress=progress)
    for key in list(sd_tv.keys()):
        res = pattern.match(key)
        if res:
            sd_tv[res.group(1) + res.group(2)] = sd_tv.pop(key)
    if model.fusion == "early" or model.stream_1_in_channels != 3:
        del sd_tv["features.conv0.weight"]
    sd = model.state_dict()
    sd.update(sd_tv)
    model.load_state_dict(sd, strict=False)
    if model.fusion == "mid":
        lidar_sd = model.stream_2_features.state_dict()
        feat_sd = model.features.state_dict()
        del feat_sd["conv0.weight"]
        lidar_sd.update(feat_sd)
        model.stream_2_features.load_state_dict(lidar_sd, strict=False)


model_urls = {
    "densenet121": "https://download.pytorch.org/models/densenet121-a639ec97.pth",
    "densenet169": "https://download.pytorch.org/models/densenet169-b2777c0a.pth",
    "densenet201": "https://download.pytorch.org/models/densenet201-c1103571.pth",
    "densenet161": "https://download.pytorch.org/models/densenet161-8d451a50.pth",
}


def _dense_u_net_lidar(arch, growth_rate, block_config, num_init_features, pretrained, progress, config):
    """(:311-332) default config when none is given; the factory overwrites the three DenseNet keys."""
    if config is None:
        config = get_config(os.path.join("content", "mnt", "My Drive", "Colab Notebooks", "DeepCV_Packages"))
    config.model.growth_rate = growth_rate
    config.model.block_config = block_config
    config.model.num_init_features = num_init_features
    model = Dense_U_Net_lidar(config)
    if pretrained:
        _load_state_dict(model, config, model_urls[arch], progress)
    return model


def densenet121_u_lidar(pretrained=False, progress=True, config=None):
    return _dense_u_net_lidar("densenet121", 32, (6, 12, 24, 16), 64, pretrained, progress, config)


def densenet161_u_lidar(pretrained=False, progress=True, config=None):
    return _dense_u_net_lidar("densenet161", 48, (6, 12, 36, 24), 96, pretrained, progress, config)


def densenet169_u_lidar(pretrained=False, progress=True, config=None):
    return _dense_u_net_lidar("densenet169", 32, (6, 12, 32, 32), 64, pretrained, progress, config)


def densenet201_u_lidar(pretrained=False, progress=True, config=None):
    return _dense_u_net_lidar("densenet201", 32, (6, 12, 48, 32), 64, pretrained, progress, config)
