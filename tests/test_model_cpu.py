"""CPU tests of the host-side mirror: module tree / state_dict contract, initialisation sequence, config
defaults, error behaviour, and that the C-ABI library loads and exports every declared symbol."""
import os
import re

import pytest
import torch

from dmmfods_b200 import _lib, config as cfgmod
from dmmfods_b200.model import Dense_U_Net_lidar, densenet121_u_lidar, densenet201_u_lidar
from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(c2=1, cb=3, **kw):
    c = cfgmod.get_config("/nonexistent")
    c.model.stream_2_in_channels = c2
    c.model.concat_before_block_num = cb
    for k, v in kw.items():
        setattr(c.model, k, v)
    return c


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dmmfods_b200.h")).read()
    declared = set(re.findall(r"\b(dmm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "symbol %s missing from libdmmfods_b200.so" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.dmm_version() >= 100


def test_struct_layouts_match_header_sizes():
    import ctypes
    from dmmfods_b200 import engine
    lib = _lib.load()
    binding = [_lib.View, _lib.Igemm, _lib.Wgrad, _lib.Bn, _lib.BnApply, _lib.BnBwd, _lib.BnBwdArgs, _lib.Head,
               _lib.HeadBwd]
    for i, st in enumerate(binding):
        assert ctypes.sizeof(st) == lib.dmm_sizeof(i), (st.__name__, ctypes.sizeof(st), lib.dmm_sizeof(i))
    assert engine._PACK_DT.itemsize == lib.dmm_sizeof(9)
    assert engine._UNPACK_DT.itemsize == lib.dmm_sizeof(10)
    assert ctypes.sizeof(_lib.GradGather) == lib.dmm_sizeof(11)
    assert engine._FOLD_DT.itemsize == lib.dmm_sizeof(12)


def test_default_config_matches_reference_defaults():
    c = cfgmod.get_config("/nonexistent")
    assert c.model.growth_rate == 32 and tuple(c.model.block_config) == (6, 12, 24, 16)
    assert c.model.concat_before_block_num == 2 and c.model.stream_2_in_channels == 1
    assert c.agent.seed == 123 and c.optimizer.learning_rate == 1e-3
    if ref_shim.reference_available():
        _, helper = ref_shim.ref_modules()
        r = helper.get_config("/nonexistent")
        for sec in ("model", "loss", "loader", "optimizer", "dataset", "agent", "scripts"):
            assert dict(c[sec]) == dict(r[sec]), sec


@pytest.mark.parametrize("c2,cb,fusion", [(0, 1, "no"), (1, 1, "early"), (1, 2, "mid"), (1, 3, "mid"), (1, 4, "mid")])
def test_fusion_modes_and_param_counts(c2, cb, fusion):
    m = densenet121_u_lidar(pretrained=False, config=_cfg(c2, cb))
    assert m.fusion == fusion
    expect = {(0, 1): 22004102, (1, 1): 22007816, (1, 2): 22409544, (1, 3): 23560136}
    if (c2, cb) in expect:
        assert m.num_params == expect[(c2, cb)]       # SURVEY Appendix A [probe]
    assert m.concat_after_module_idx == 3 + 2 * (cb - 1)


def test_invalid_fusion_raises_attribute_error():
    with pytest.raises(AttributeError):
        Dense_U_Net_lidar(_cfg(1, 5))
    with pytest.raises(AttributeError):
        Dense_U_Net_lidar(_cfg(0, 0))


def test_cpu_call_fails_loudly():
    m = densenet121_u_lidar(pretrained=False, config=_cfg(0, 1))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 96), torch.zeros(1, 1, 64, 96))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("c2,cb", [(0, 1), (1, 1), (1, 2), (1, 3)])
def test_state_dict_and_init_identical_to_reference(c2, cb):
    model_mod, _ = ref_shim.ref_modules()
    torch.manual_seed(123)
    ref = model_mod.densenet121_u_lidar(pretrained=False, config=ref_shim.ref_config(c2, cb))
    torch.manual_seed(123)
    ours = densenet121_u_lidar(pretrained=False, config=_cfg(c2, cb))
    sr, so = ref.state_dict(), ours.state_dict()
    assert list(sr.keys()) == list(so.keys())
    for k in sr:
        assert sr[k].shape == so[k].shape and sr[k].dtype == so[k].dtype, k
        assert torch.equal(sr[k], so[k]), "initial value of %s differs" % k
    ours.load_state_dict(sr, strict=True)
    ref.load_state_dict(so, strict=True)
    assert [k for k, _ in ref.named_parameters()] == [k for k, _ in ours.named_parameters()]


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_densenet201_keys_match_reference():
    model_mod, _ = ref_shim.ref_modules()
    with torch.device("meta"):
        ref = model_mod.densenet201_u_lidar(pretrained=False, config=ref_shim.ref_config(1, 3))
        ours = densenet201_u_lidar(pretrained=False, config=_cfg(1, 3))
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert ours.num_params == 57346504


@pytest.mark.parametrize("c2,cb", [(0, 1), (1, 1), (1, 2), (1, 3), (1, 4)])
def test_engine_plan_covers_every_parameter(c2, cb):
    """host logic without a GPU: the launch programs are built for every fusion mode, every convolution weight
    has a pack job (forward + data gradient) and an unpack job (weight gradient), every BatchNorm a backward."""
    from dmmfods_b200.engine import Engine
    c = _cfg(c2, cb, growth_rate=16, block_config=(2, 4, 2, 2), num_init_features=32, bn_size=2)
    m = Dense_U_Net_lidar(c)
    params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in m.state_dict(keep_vars=True).items()}
    eng = Engine(params, m.model_cfg(), 2, 64, 96, plan_only=True)
    conv_w = {k for k, v in params.items() if v.dim() == 4}
    bn_w = {k for k, v in params.items() if v.dim() == 1 and k.endswith(".weight")}
    assert {j["wname"] for j in eng._unpack_jobs} == conv_w
    names = [op.name for op in eng.bwd]
    for k in bn_w:
        pre = k[:-len(".weight")]
        assert any(n.startswith(pre + ".bwd") for n in names), pre
    assert len({op.name for op in eng.fwd}) == len(eng.fwd)
    # first writer of each block-gradient buffer stores, later ones accumulate
    seen = {}
    for op in eng.bwd:
        if op.gbuf is not None:
            assert op.arg.out_mode == (2 if id(op.gbuf) in seen else 1)
            seen[id(op.gbuf)] = True
    with pytest.raises(RuntimeError):
        eng.forward(torch.zeros(2, 3, 64, 96), torch.zeros(2, 1, 64, 96))
    # gradient buckets: contiguous, ordered, covering the whole flat gradient buffer, in backward order
    small = Engine(params, m.model_cfg(), 2, 64, 96, plan_only=True, bucket_bytes=100 << 10)
    assert len(small.segments) > 2
    pos = 0
    for ops_, job_lo, njobs, lo, hi in small.segments:
        assert lo == pos and hi >= lo
        pos = hi
    assert pos == small.gflat.numel() == sum(p.numel() for p in m.parameters())
    assert sum(s[2] for s in small.segments) == len(small._unpack_jobs)
    assert small.param_names[0].startswith("dec_out_to_heat_maps")      # the head finishes first in backward
    assert Engine.gradient_order(params, m.model_cfg(), 2, 64, 96) == eng.param_names


def test_engine_rejects_sizes_the_reference_rejects():
    from dmmfods_b200.engine import Engine
    c = _cfg(1, 3, growth_rate=16, block_config=(2, 2, 2, 2), num_init_features=32, bn_size=2)
    m = Dense_U_Net_lidar(c)
    params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in m.state_dict(keep_vars=True).items()}
    with pytest.raises(ValueError):
        Engine(params, m.model_cfg(), 1, 72, 96, plan_only=True)     # 72/4 = 18 -> 9 is odd at block 2
    with pytest.raises(ValueError):
        Engine(params, m.model_cfg(), 1, 63, 96, plan_only=True)


def test_row_phase_views_cover_all_rows():
    """ops.Mat.row_phase_view: the even / odd image-row views used by the stem's 7-tap vertical convolution (input row
    2*oy + kh - 3 = 2*(oy + dy) + p) - pure host geometry, checked on a CPU tensor."""
    import torch
    from dmmfods_b200 import ops
    B, H, W, ld = 2, 7, 5, 16
    t = torch.arange(B * H * W * ld, dtype=torch.float32).to(torch.bfloat16).view(B * H * W, ld)
    m = ops.Mat(t, B, H, W)
    ev, od = m.row_phase_view(0), m.row_phase_view(1)
    assert (ev.H, od.H) == ((H + 1) // 2, H // 2) and ev.W == od.W == W and ev.C == od.C == ld
    assert ev.sw == ld and ev.sh == 2 * W * ld and ev.sb == H * W * ld
    assert od.ptr - ev.ptr == 2 * W * ld                      # one image row further, in bytes of bf16
    taps = [((kh - 3) % 2, (kh - 3 - (kh - 3) % 2) // 2) for kh in range(7)]
    assert [2 * dy + p for p, dy in taps] == [kh - 3 for kh in range(7)]


def test_default_plan_uses_the_folded_and_unfolded_formulations():
    """structure of the DenseNet-121 mid-fusion plan built without a GPU: the formulations DESIGN.md describes are the ones the
    engine emits by default (growth convolutions with folded kernel columns, 5-tap head convolution with fp32 NCHW output,
    stem unfold for the RGB stream / im2col for the LiDAR stream, conv1 with the BN-ReLU prologue, KxK prologue off)."""
    from dmmfods_b200.engine import Engine
    m = Dense_U_Net_lidar(_cfg(1, 3))
    params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in m.state_dict(keep_vars=True).items()}
    eng = Engine(params, m.model_cfg(), 2, 64, 96, plan_only=True)
    fwd = {op.name: op for op in eng.fwd}
    conv2 = [op for n, op in fwd.items() if n.endswith(".conv2")]
    assert len(conv2) == 2 * (6 + 12) + 24 + 16
    assert all(op.arg.out_mode == 3 and op.arg.fold_kw == 3 and op.arg.N == 96 and op.arg.num_taps == 3 and not op.arg.pro_enable
               for op in conv2)
    conv1 = [op for n, op in fwd.items() if n.endswith(".conv1")]
    assert len(conv1) == len(conv2) and all(op.arg.pro_enable == 1 and op.arg.num_taps == 1 for op in conv1)
    r1 = fwd["dec_out_to_heat_maps.refine1"].arg
    assert (r1.out_mode, r1.fold_kw, r1.N, r1.num_taps) == (2, 5, 15, 5)
    r0 = fwd["dec_out_to_heat_maps.refine0"].arg
    assert (r0.out_mode, r0.num_taps, r0.N) == (0, 9, 64)
    s1, s2 = fwd["features.conv0"].arg, fwd["stream_2_features.conv0"].arg
    assert (s1.num_src, s1.num_taps) == (2, 7) and (s2.num_src, s2.num_taps) == (1, 1)
    # backward: refine1's data gradient is the 5-tap vertical form, its weight gradient one launch over 5 taps
    bwd = {op.name: op for op in eng.bwd}
    assert bwd["dec_out_to_heat_maps.refine1.dgrad"].arg.num_taps == 5
    assert bwd["dec_out_to_heat_maps.refine1.wgrad"].arg.num_b == 5
    assert any(n.endswith(".norm2") for n in fwd)                      # norm2 is still a separate BN-ReLU pass (KxK prologue off)


@pytest.mark.parametrize("factory", ["densenet121_u_lidar", "densenet161_u_lidar", "densenet169_u_lidar", "densenet201_u_lidar"])
def test_training_plans_of_all_four_factories(factory):
    """every exported factory (Dense_U_Net_lidar.py:335-388) yields a TRAINING engine plan: densenet201's 48-layer block has
    50 consumers of its first channels, more than one dmm_grad_gather launch took in round 1 (DMM_GATHER_MAX)."""
    from dmmfods_b200 import _lib as L, model as M
    from dmmfods_b200.engine import Engine
    m = getattr(M, factory)(pretrained=False, config=_cfg(1, 3))
    params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in m.state_dict(keep_vars=True).items()}
    eng = Engine(params, m.model_cfg(), 1, 64, 96, plan_only=True)
    assert sorted(eng.param_names) == sorted(k for k, _ in m.named_parameters())
    gathers = [g for g in eng._gathers]
    assert gathers and all(1 <= len(g[5]) for g in gathers)
    most = max(sum(d.nsrc for d in g[5]) - (len(g[5]) - 1) for g in gathers)
    assert most == max(m.block_config) + 2
    assert all(d.nsrc <= L.GATHER_MAX and d.nk <= L.GATHER_MAX for g in gathers for d in g[5])


def test_gather_chain_beyond_gather_max(monkeypatch):
    """more consumers than DMM_GATHER_MAX: the gather is emitted as a chain of launches that re-read their own output."""
    from dmmfods_b200 import _lib as L
    from dmmfods_b200.engine import Engine
    monkeypatch.setattr(L, "GATHER_MAX", 8)
    c = _cfg(1, 3, growth_rate=16, block_config=(2, 12, 2, 2), num_init_features=32, bn_size=2)
    m = Dense_U_Net_lidar(c)
    params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in m.state_dict(keep_vars=True).items()}
    eng = Engine(params, m.model_cfg(), 1, 64, 96, plan_only=True)
    chained = [g for g in eng._gathers if len(g[5]) > 1]
    assert chained
    for g in chained:
        descs, dst = g[5], g[4]
        assert all(d.nsrc <= 8 for d in descs)
        assert all(d.src[0] == dst.ptr().value and d.plane[0] == 0 for d in descs[1:])      # accumulator first
        assert sum(d.nsrc for d in descs) - (len(descs) - 1) == len([c_ for c_ in g[1].contribs if c_["C"] >= g[2] + g[3]])
    assert eng.gather_launches > len(eng._gathers)


def test_inference_plan_folds_batchnorms_into_their_producers():
    """eval-mode engines: norm2 -> conv1, decoder norm1 -> conv_reduce, head norm1 -> refine0 are folded (scaled
    packed rows + bias / ReLU epilogue); what is left are the BatchNorms over multi-consumer block buffers and the pooled ones."""
    from dmmfods_b200.engine import Engine
    m = Dense_U_Net_lidar(_cfg(1, 3))
    params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in m.state_dict(keep_vars=True).items()}
    ev = Engine(params, m.model_cfg(), 1, 64, 96, plan_only=True, training=False, need_backward=False)
    tr = Engine(params, m.model_cfg(), 1, 64, 96, plan_only=True, training=True, need_backward=False)
    n_layers = 2 * (6 + 12) + 24 + 16
    assert len(ev._fold_jobs) == n_layers + 4 + 1            # dense layers + decoder stages + head
    assert len(tr._fold_jobs) == 0
    assert len(ev.fwd) == len(tr.fwd) - len(ev._fold_jobs) - 1          # (- the input-statistics launch of training mode)
    folded = [op for op in ev.fwd if op.kind == "igemm_fprop" and op.arg.epi_bias]
    assert len(folded) == len(ev._fold_jobs) and all(op.arg.epi_relu == 1 and op.arg.out_mode == 0 for op in folded)
    conv1 = [op for op in ev.fwd if op.name.endswith(".conv1")]
    assert all(op.arg.pro_enable == 1 and op.arg.epi_bias for op in conv1)       # norm1 as prologue, norm2 as epilogue
    assert not any(op.name.endswith(".norm2") for op in ev.fwd)
    left = sorted(op.name for op in ev.fwd if op.kind == "bn_relu_apply")
    assert len(left) == 16 and all(("norm0" in n) or ("norm+pool" in n) or ("concat_module" in n) for n in left), left


@pytest.mark.parametrize("precision,dtype", [("tf32", 1), ("tf32x3", 2)])
def test_strict_plan(precision, dtype):
    """the strict forward plan: fp32 buffers, every convolution an unfused / unfolded dtype-1 (tf32) or dtype-2 (3xTF32) launch."""
    from dmmfods_b200.engine import Engine
    m = Dense_U_Net_lidar(_cfg(1, 3))
    params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in m.state_dict(keep_vars=True).items()}
    eng = Engine(params, m.model_cfg(), 1, 64, 96, plan_only=True, need_backward=False, precision=precision)
    convs = [op for op in eng.fwd if op.kind == "igemm_fprop"]
    assert convs and all(op.arg.dtype == dtype and op.arg.kwidth == 32 and op.arg.out_mode in (0, 1) and not op.arg.pro_enable
                         and not op.arg.fold_kw for op in convs)
    from dmmfods_b200.ops import Mat
    assert eng._wpk.dtype == torch.float32 and all(m_.t.dtype == torch.float32 for m_ in eng._keep if isinstance(m_, Mat))
    if dtype == 2:          # [values ; remainders]: twice the packed weights
        single = Engine(params, m.model_cfg(), 1, 64, 96, plan_only=True, need_backward=False, precision="tf32")
        assert eng._wpk.numel() == 2 * single._wpk.numel()
    with pytest.raises(NotImplementedError):
        Engine(params, m.model_cfg(), 1, 64, 96, plan_only=True, need_backward=True, precision=precision)


def test_work_table_covers_every_element_once():
    from dmmfods_b200.engine import WORK_CHUNK, _work_table
    sizes = [1, WORK_CHUNK, WORK_CHUNK + 1, 5 * WORK_CHUNK - 3, 0, 64]
    t = _work_table(sizes, "cpu")
    assert t.dtype == torch.int32 and t.shape[1] == 2
    for j, n in enumerate(sizes):
        chunks = sorted(int(c) for jj, c in t.tolist() if jj == j)
        assert chunks == list(range((n + WORK_CHUNK - 1) // WORK_CHUNK))


def _view(C_, W, H, B):
    v = _lib.View()
    v.ptr, v.C, v.W, v.H, v.B = 4096, C_, W, H, B
    v.sw, v.sh, v.sb = C_, C_ * W, C_ * W * H
    return v


def test_convtranspose_wgrad_plan_groups_taps_per_phase():
    """ConvTranspose2d(128, 128, 3, stride 2): ONE weight-gradient launch per output phase (1 / 2 / 2 / 4 taps as A chunks at their
    shifts, the phase view once as B) instead of one launch per tap; every (tap, input channel) row of the scratch matrix is
    written by exactly one A chunk."""
    from dmmfods_b200 import ops
    x, ys = _view(128, 240, 160, 32), [_view(128, 240, 160, 32) for _ in range(4)]
    taps, _ = ops.convt_wgrad_taps()
    plan = ops.plan_conv_wgrad(x, ys, taps, 128, 128)
    assert len(plan["launches"]) == 4 and plan["tap_order"] == list(range(9))
    assert (plan["rows"], plan["ld"], plan["dt"], plan["dm"], plan["dn"]) == (9 * 128, 128, 128 * 128, 128, 1)
    rows = []
    for l in plan["launches"]:
        (ysrc, dy, dx, ch0, out0), = l["b_slots"]
        assert (dy, dx, ch0, out0) == (0, 0, 0, 0)
        for (s_, ady, adx, c0, o0) in l["a_slots"]:
            t = o0 // 128
            assert s_ == 0 and taps[t] == (ysrc, ady, adx) and o0 % 128 == c0 and c0 in (0, 64)
            rows.append(o0)
        d = ops.make_wgrad(W=240, H=160, B=32, dw=0, ld=plan["ld"], **l)
        assert ((d.num_a + 1) // 2) * d.num_b * d.n_tile <= 512          # TMEM columns
    assert sorted(rows) == [64 * i for i in range(18)]
    # 256 channels: an accumulator per tap already fills TMEM - one launch per tap as before
    plan = ops.plan_conv_wgrad(_view(256, 120, 80, 32), [_view(256, 120, 80, 32) for _ in range(4)], taps, 256, 256)
    assert len(plan["launches"]) == 9


def test_wgrad_plan_uses_one_m_tile_per_cta_for_small_layers(monkeypatch):
    from dmmfods_b200 import ops
    monkeypatch.delenv("DMM_WGRAD_NA", raising=False)
    taps = ops.conv_taps(1, 0)[0]
    small = ops.plan_conv_wgrad(_view(1024, 30, 20, 32), [_view(128, 30, 20, 32)], taps, 1024, 128)["launches"][0]
    big = ops.plan_conv_wgrad(_view(256, 240, 160, 32), [_view(128, 240, 160, 32)], taps, 256, 128)["launches"][0]
    assert len(small["a_slots"]) == 2 and small["ya"] == 8 and small["a_step"] == 128          # 19 200 pixels: 128 channels per CTA
    assert len(big["a_slots"]) == 4 and big["ya"] == 1                                          # 1.2 M pixels: all 256 channels in one CTA


def test_family_wgrad_swaps_image_axes_when_that_pads_less():
    """3x3 weight gradient with 32 output-gradient channels (family launch, 8 x 16 pixel tiles): on a 30 x 20 image the views and
    shifts are handed to the kernel with x and y exchanged (24 x 32 instead of 32 x 32 padded pixels); 240 x 160 stays as it is."""
    from dmmfods_b200 import ops
    taps = ops.conv_taps(3, 1)[0]
    for (W, H, swapped) in ((30, 20, True), (240, 160, False), (60, 40, True)):
        x, y = _view(128, W, H, 2), _view(32, W, H, 2)
        l = ops.plan_conv_wgrad(x, [y], taps, 128, 32)["launches"][0]
        d = ops.make_wgrad(W=W, H=H, B=2, dw=0, ld=9 * 32, **l)
        assert d.tile_w == 8 and (d.W, d.H) == ((H, W) if swapped else (W, H))
        assert (d.b_src[0].W, d.b_src[0].sw) == ((H, 32 * W) if swapped else (W, 32))
        for i, (s_, dy, dx, c0, o0) in enumerate(l["b_slots"]):
            assert (d.b[i].dy, d.b[i].dx) == ((dx, dy) if swapped else (dy, dx))
