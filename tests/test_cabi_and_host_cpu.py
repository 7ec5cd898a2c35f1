"""CPU tier: the C-ABI library loads without a GPU and exports exactly what include/xcp.h declares; the Python
binding table matches the header; the drop-in modules reproduce the reference's state_dict schema and seeded
init; the product never imports the oracle; CPU inputs fail loudly (no fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from multimodal_deepfake_detection_b200 import _lib  # noqa: E402


def _header_decls():
    src = open(os.path.join(ROOT, "include", "xcp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(xcp_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.groups()
        sig = ""
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                sig += "p" if "*" in a else {"long long": "l", "int": "i", "float": "f", "double": "d"}[
                    "long long" if a.startswith("long long") else a.split()[0]]
        out[name] = sig
    return out


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.isfile(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "multimodal_deepfake_detection_b200", "csrc"), "-j", "8"])
    return ctypes.CDLL(_lib.LIB_PATH)


def test_library_exports_every_header_symbol(built_lib):
    decls = _header_decls()
    assert len(decls) >= 40
    for name in decls:
        assert hasattr(built_lib, name), "libxcp_sm100.so does not export %s" % name
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r" T (xcp_\w+)", out))
    assert exported == set(decls), (exported ^ set(decls))


def test_binding_table_matches_header():
    decls = _header_decls()
    decls.pop("xcp_last_error_string")
    assert decls == _lib.SIGNATURES


def test_host_only_calls_work_without_gpu(built_lib):
    assert _lib.call("xcp_version") >= 100
    if not torch.cuda.is_available():
        # a compute call without a device must return an error code and a message, not crash
        with pytest.raises(_lib.XcpError):
            _lib.call("xcp_check_device", 0)


def test_no_libcuda_or_torch_link_dependency():
    out = subprocess.check_output(["ldd", _lib.LIB_PATH]).decode()
    assert "libcuda.so" not in out and "libtorch" not in out and "libc10" not in out


def test_sass_is_blackwell_native():
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "FFMA2"):     # tcgen05.mma, TMA load, tcgen05.ld, packed fp32 FMA
        assert mnemonic in out, mnemonic
    assert "HGMMA" not in out
    # the cluster LSTM kernels: cluster-wide barriers around the DSMEM h / dh exchange, W_hh fragments fed to HMMA from registers
    for mnemonic in ("UCGABAR_ARV", "UCGABAR_WAIT", "HMMA.16816"):
        assert mnemonic in out, mnemonic


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodal_deepfake_detection_b200")
    offenders = []
    for d in (pkg, os.path.join(ROOT, "Models"), os.path.join(ROOT, "Dataset")):
        for base, _, files in os.walk(d):
            for f in files:
                if f.endswith(".py"):
                    txt = open(os.path.join(base, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "xception_oracle" in txt or "mfcc_oracle" in txt:
                        offenders.append(os.path.join(base, f))
    for f in ("train_visual.py", "train_audio.py", "train_au_face.py", "train_au_patch.py", "test_visual.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p) and re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), flags=re.M):
            offenders.append(p)
    assert not offenders, offenders


def test_state_dict_schema_and_seeded_init_match_reference_restatement():
    from oracle import xception_oracle as O
    from Models.Xception import Xception
    from Models.XceptionLSTMV import XceptionLSTMV
    torch.manual_seed(3)
    net = Xception(num_classes=2)
    sd = net.state_dict()
    shapes = O.xception_param_shapes(2)
    assert [k for k, _, _ in shapes] == list(sd.keys())
    assert all(tuple(sd[k].shape) == tuple(s) for k, s, _ in shapes)
    assert len(sd) == 276
    assert sd["bn1.num_batches_tracked"].dtype == torch.long
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = XceptionLSTMV(16)
    keys = list(m.state_dict().keys())
    assert len(keys) == 288
    assert keys[-14:] == ["lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
                          "fc_layers.0.weight", "fc_layers.0.bias", "fc_layers.3.weight", "fc_layers.3.bias",
                          "fc_layers.6.weight", "fc_layers.6.bias", "fc_layers.9.weight", "fc_layers.9.bias",
                          "fc_out.weight", "fc_out.bias"]
    assert not any(p.requires_grad for p in m.feature_extractor.parameters())      # XceptionLSTMV.py:15-16
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 4 * 16 * 2048 + 4 * 16 * 16 + 8 * 16 + \
        16 * 1024 + 1024 + 3 * (1024 * 1024 + 1024) + 1024 + 1


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted (GPU box)")
def test_seeded_init_bit_identical_to_live_reference():
    from oracle import ref_shim
    from Models.Xception import Xception
    ref = ref_shim.load()
    torch.manual_seed(11); a = Xception(num_classes=5)
    torch.manual_seed(11); b = ref.Xception(num_classes=5)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
    assert [n for n, _ in a.named_modules()] == [n for n, _ in b.named_modules()]


def test_cpu_inputs_fail_loudly():
    from multimodal_deepfake_detection_b200 import SeparableConv2d, Xception, XcpError
    with pytest.raises(XcpError):
        Xception(num_classes=2)(torch.zeros(1, 3, 75, 75))
    with pytest.raises(XcpError):
        SeparableConv2d(8, 8, 3, 1, 1)(torch.zeros(1, 8, 5, 5))


def test_block_spec_matches_reference_layout():
    from multimodal_deepfake_detection_b200 import Block
    from oracle import xception_oracle as O
    for (name, cin, cout, reps, stride, swr, gf) in O.BLOCKS:
        blk = Block(cin, cout, reps, stride, start_with_relu=swr, grow_first=gf)
        spec = blk._spec()
        want = O.block_layout(cin, cout, reps, swr, gf)
        assert [(u.cin, u.cout, u.relu) for u in spec.units] == [(ci, co, r) for (_, _, ci, co, r) in want]
        keys = [k for k in blk.state_dict() if k.endswith("conv1.weight")]
        assert keys == ["rep.%d.conv1.weight" % i for (i, _, _, _, _) in want]


def test_loop_metrics_match_sklearn():
    """binary_metrics / youden_threshold (loops.py) reproduce the sklearn calls of train_visual.py:476-487."""
    np = pytest.importorskip("numpy")
    skm = pytest.importorskip("sklearn.metrics")
    from multimodal_deepfake_detection_b200.loops import binary_metrics, youden_threshold
    rng = np.random.default_rng(0)
    for n, ties in [(200, True), (57, False), (1000, True)]:
        y = rng.integers(0, 2, n)
        p = np.clip(y * 0.3 + rng.random(n) * 0.8, 0, 1)
        if ties:
            p = np.round(p, 2)
        m = binary_metrics(y, p)
        fpr, tpr, thr = skm.roc_curve(y, p)
        assert abs(m["AUC"] - skm.roc_auc_score(y, p)) < 1e-9
        assert abs(m["AP"] - skm.average_precision_score(y, p)) < 1e-9
        fnr = 1 - tpr
        k = np.nanargmin(np.abs(fpr - fnr))
        assert abs(m["EER"] - (fpr[k] + fnr[k]) / 2) < 1e-9
        assert abs(m["pAUC"] - skm.auc(fpr[fpr <= 0.1], tpr[fpr <= 0.1]) / 0.1) < 1e-9
        t, f_, t_ = youden_threshold(y, p)
        assert abs((t_ - f_) - np.max(tpr - fpr)) < 1e-12
    assert binary_metrics([1, 1, 1], [0.2, 0.3, 0.9])["EER"] == 1.0      # single-class sentinel of the reference


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line with the metric / unit of the product arm, the
    `cpu_baseline` description of the run and a zero-copy `e2e`; it must work without a GPU."""
    import json
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("train clips/sec XceptionLSTMV") and d["unit"] == "clips/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] and d["vs_baseline"] is None


def test_bench_kernel_roofline_decoder_uses_logical_channels():
    """bench.py turns the (entry point, argument signature) log of _lib.timer_* into per-family rooflines: algorithmic work is
    counted with the LOGICAL 728 channels (the 768 pitch inflated the fraction by 11 %), mixed shapes weigh by time."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    M = 92416
    fam, flops, nbytes = bench._work("xcp_gemm_tn", (True, 768, True, 768, True, 768, M, 768, 768, 1, True, False, 0, True))
    assert "fwd" in fam and flops == 2.0 * M * 728 * 728 and nbytes == 2.0 * M * 728 * 2 + 2.0 * 728 * 728
    fam, flops, nbytes = bench._work("xcp_dw3x3_bwd", (True, True, True, True, True, 1, True, True, False, True, True, 256, 19, 19, 768, 728, 0, True))
    assert fam.endswith("19x19") and nbytes == 8.0 * 256 * 19 * 19 * 728          # dD + x + dz + the identity-skip gradient
    fam, _, nbytes = bench._work("xcp_bn_bwd", (0, True, True, False, False, True, True, True, True, True, 1, True, False, True, True, True, True,
                                               256, 19, 19, 768, 728, 0, 0, 0, True))
    assert nbytes == 2.0 * 256 * 19 * 19 * 728 * 3                               # presums given: only the apply pass (y, dz -> dy)
    assert bench._work("xcp_lstm_fwd", ()) is None
    peaks = {"hbm_gbs": 6450.6, "bf16_tflops_sustained": 1428.5}
    log = {("xcp_gemm_tn", (True, 768, True, 768, True, 768, M, 768, 768, 1, True, False, 0, True)): [0.0903 * 2, 2],
           ("xcp_lstm_fwd", ()): [0.05, 1]}
    ks, total = bench.kernel_rooflines(log, 1.0, peaks, 1)
    top = ks[0]
    assert top["bound"] == "tensor" and abs(top["frac"] - (2.0 * M * 728 * 728 / 90.3e-6 / 1428.5e12)) < 1e-3 and 0.75 < top["frac"] < 0.77
    assert ks[1]["bound"] == "latency" and ks[1]["frac"] is None and abs(sum(k["share_of_step"] for k in ks) - 1.0) < 1e-9


def test_staged_reference_is_byte_identical_and_untracked():
    """bench.py's reference arm runs the UNMODIFIED reference files staged by build() under the git-ignored baseline/_ref."""
    import __graft_entry__ as ge
    if not ge.stage_reference():
        pytest.skip("reference tree not mounted")
    for f in ge.REF_FILES:
        a = open(os.path.join(ge.REF_DIR, f), "rb").read()
        b = open(os.path.join(ROOT, "baseline", "_ref", "RefModels", f), "rb").read()
        assert a == b
    out = subprocess.run(["git", "check-ignore", "baseline/_ref/RefModels/Xception.py"], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0
