"""Host-side behaviour that needs no GPU: plugin names and signatures mirror the reference, CPU
tensors are refused (no fallback), the drop-in path glue resolves reference-only modules."""
import inspect
import os
import subprocess
import sys

import pytest
import torch

from conftest import PKG, ROOT

REF = "/root/reference"


def test_plugin_names_and_signatures():
    from losses.centernet import DetectionLoss
    from losses.entropy import EntropyLoss
    from losses.max_square import MaxSquareLoss
    from losses.advent import AdventLoss
    from backends.decode import decode_detection
    from utils.image import entropy_map
    from utils.tensor import _sigmoid, _gather_feat, _transpose_and_gather_feat   # noqa: F401
    sig = inspect.signature(DetectionLoss.__init__)
    # the reference's kwargs, in its order (losses/centernet.py:8-11), then this package's one extension
    assert list(sig.parameters)[1:] == ["hm_weight", "wh_weight", "off_weight", "kp_weight", "angle_weight",
                                        "periodic", "kp_indices", "kp_distance_weight", "kp_distance_weight_l1",
                                        "max_detections"]
    assert sig.parameters["max_detections"].default is None
    assert sig.parameters["angle_weight"].default == 1.0 and sig.parameters["periodic"].default is False
    assert list(inspect.signature(decode_detection).parameters) == ["heat", "wh", "reg", "kps", "K", "rotated",
                                                                    "nms_size"]
    assert inspect.signature(decode_detection).parameters["K"].default == 100
    assert list(inspect.signature(EntropyLoss.__init__).parameters) == ["self", "eta"]
    for cls in (DetectionLoss, EntropyLoss, MaxSquareLoss, AdventLoss):
        assert issubclass(cls, torch.nn.Module)
    assert callable(entropy_map)


def test_cpu_tensors_are_refused():
    from losses.entropy import EntropyLoss
    from losses.max_square import MaxSquareLoss
    from losses.advent import AdventLoss
    from backends.decode import decode_detection
    from utils.image import entropy_map
    x = torch.zeros(1, 3, 4, 4)
    for call in (lambda: EntropyLoss()({"hm": x}, None), lambda: MaxSquareLoss()({"hm": x}, None),
                 lambda: AdventLoss()(x, 1), lambda: entropy_map(x),
                 lambda: decode_detection(x, x[:, :2], x[:, :2], K=4)):
        with pytest.raises(RuntimeError, match="CUDA only"):
            call()


def test_decode_rejects_other_nms_windows():
    from backends.decode import decode_detection
    with pytest.raises(NotImplementedError):
        decode_detection(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4), nms_size=5)


def test_product_code_never_imports_the_oracle():
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(root, f)


def test_synthetic_inputs_are_deterministic_and_sliceable():
    from cnhead import synthetic
    cfg = synthetic.CONFIGS["cfg2"]
    a = synthetic.make_inputs(cfg, batch=4)
    b = synthetic.make_inputs(cfg, batch=2, sample_offset=2)
    for grp in ("output", "batch"):
        for k in a[grp]:
            assert torch.equal(a[grp][k][2:4], b[grp][k]), (grp, k)
    gt = a["batch"]["hm"]
    assert float(gt.max()) == 1.0 and float(gt.min()) == 0.0
    n_obj = a["batch"]["reg_mask"].sum(1)
    assert (n_obj >= 1).all() and (n_obj <= 20).all()
    assert synthetic.CONFIGS["cfg2"].bytes_per_sample() == 2228224
    assert synthetic.CONFIGS["cfg3"].bytes_per_sample() == 2293760
    assert synthetic.CONFIGS["cfg5"].bytes_per_sample() == 26476544
    assert synthetic.CONFIGS["cfg4"].bytes_per_sample(decode=False) == 3407872


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")
def test_dropin_shadowing_of_a_reference_checkout():
    """PYTHONPATH=<ours>:<reference>: our modules win, reference-only modules and names still resolve."""
    code = (
        "import losses.centernet as c, backends.decode as d, utils.image as i, utils.box as b;"
        "assert 'centernet-uda_b200' in c.__file__ and 'centernet-uda_b200' in d.__file__;"
        "assert b.__file__.startswith('%s');"                      # reference-only module
        "assert callable(i.gaussian_radius) and callable(i.draw_umich_gaussian);"   # re-exported names
        "assert hasattr(c, 'FocalLoss') and hasattr(d, '_nms');"
        "print('ok')" % REF)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG, REF]))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_host_feeder_and_fetcher_refuse_to_run_without_cuda():
    """the staging helpers are plumbing around the CUDA path: no CPU mode"""
    from cnhead.feeder import HostFeeder
    with pytest.raises(RuntimeError, match="CUDA"):
        HostFeeder("cpu")
    with pytest.raises(ValueError):
        HostFeeder("cuda", depth=0)


def test_graphed_step_refuses_to_run_without_cuda():
    from cnhead.graphed import GraphedFn, HostStep
    with pytest.raises(RuntimeError):
        GraphedFn(lambda: {}, "cpu")
    with pytest.raises(RuntimeError):
        HostStep(lambda: {}, "cpu")


def test_host_feeder_span_detection():
    """HostFeeder ships a set as one copy only when it is a dense, aligned run of slices of ONE storage."""
    from cnhead.feeder import HostFeeder
    buf = torch.empty(1 << 16, dtype=torch.uint8)

    def carve(t, off):
        n = t.numel() * t.element_size()
        out = buf[off:off + n].view(t.dtype).view(t.shape)
        out.copy_(t)
        return out

    a, b, c = torch.randn(2, 3, 8, 8), torch.randint(0, 64, (2, 5)), torch.ones(2, 5, dtype=torch.uint8)
    ha, hb, hc = carve(a, 4096), carve(b, 8192), carve(c, 12288)
    lo, span, layout = HostFeeder._span(({"hm": ha}, {"ind": hb, "mask": hc}))
    assert lo == 4096 and span == 8192 + 10
    assert layout == ((0, torch.float32, (2, 3, 8, 8)), (4096, torch.int64, (2, 5)), (8192, torch.uint8, (2, 5)))
    # separate storages, a misaligned slice, or a sparse span: per-tensor copies instead
    assert HostFeeder._span(({"hm": a}, {"ind": b})) is None
    assert HostFeeder._span(({"hm": ha}, {"mask": buf[12289:12299]})) is None
    big = torch.empty(1 << 22, dtype=torch.uint8)
    assert HostFeeder._span(({"x": big[:16]}, {"y": big[(1 << 22) - 16:]})) is None
    assert HostFeeder._span(({"hm": ha.transpose(2, 3)},)) is None


def test_pinned_arena_layout_arithmetic():
    from cnhead.feeder import PinnedArena
    ts = [torch.empty(3, 5), torch.empty(7, dtype=torch.int64), torch.empty(1, dtype=torch.uint8)]
    assert PinnedArena.bytes_for(ts) == 3 * 4096 and PinnedArena.MIN_BYTES >= 256 << 20


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys,
    the same `config` dict the GPU arm prints, `--warmup` honoured, and no CUDA needed."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "heatmaps/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 3 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    _sys.path.insert(0, root)
    import bench
    from cnhead import synthetic
    assert line["config"] == bench.config_dict(synthetic.CONFIGS["cfg2"], 16, 1)


def test_bench_gpu_arm_fails_loudly_without_cuda():
    import subprocess
    import sys as _sys
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "CUDA" in r.stderr
