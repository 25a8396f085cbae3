"""CPU: host-side logic - geometry tables, the C-ABI library's exports, loud failure without a
GPU, pool sharding and the world_size-2 gather (gloo)."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from aod_meh_hua_b200 import _lib
from aod_meh_hua_b200.pool import shard_range
from aod_meh_hua_b200.specs import SPECS, ScoringParams, get_spec, parse_agg_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_survey_shape_table():
    # N, K_tot and algorithmic bytes/image per BASELINE config
    want = {
        "cfg1_retina_r50_512_voc": (49104, 3720, 4.556e6),
        "cfg2_ssd300_voc": (8732, 2790, 1.103e6),
        "cfg3_retina_r50_800x1344_coco": (201600, 4693, 66.989e6),
        "cfg3p_retina_r50_800x800_coco": (120087, 4441, 40.489e6),
        "cfg4_ssd512_coco": (24564, 3500, 9.317e6),
        "cfg5_retina_r101_1344_coco": (338454, 5000, 111.439e6),
    }
    for name, (n, k, nbytes) in want.items():
        sp = get_spec(name)
        assert sp.num_priors == n and sp.k_tot == k
        assert abs(sp.k1_bytes_per_image() - nbytes) / nbytes < 1e-3


def test_header_symbols_are_exported_and_bound():
    """Every function include/mehhua.h declares is exported by libmehhua.so and bound in _lib."""
    hdr = open(os.path.join(ROOT, "include", "mehhua.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mehhua_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mehhua_abi_version() == _lib.ABI_VERSION
    assert lib.mehhua_launch_count() >= 0


def test_struct_layouts_match_header():
    lib = _lib.load()
    # geometry-only entry points run without a GPU: rows per image and workspace size
    from aod_meh_hua_b200.scoring import make_config
    for name in ("cfg1_retina_r50_512_voc", "cfg2_ssd300_voc", "cfg3_retina_r50_800x1344_coco", "cfg4_ssd512_coco"):
        sp = get_spec(name)
        cfg = make_config(sp, ScoringParams(), 4096)
        lv = _lib.LevelArray()
        for s, ((h, w), a) in enumerate(zip(sp.featmaps, sp.num_anchors)):
            lv[s].H, lv[s].W, lv[s].A = h, w, a
        assert lib.mehhua_rows_per_image(C.byref(cfg), lv) == sp.k_tot
        ws = lib.mehhua_workspace_bytes(C.byref(cfg), lv, 4)
        num_fg = sp.num_classes
        assert ws >= 4 * (4 * sp.num_priors + 8 * sp.k_tot * num_fg)
    bad = make_config(get_spec("cfg2_ssd300_voc"), ScoringParams(), 4096)
    bad.max_per_img = 10_000
    assert lib.mehhua_workspace_bytes(C.byref(bad), lv, 4) == 0
    assert b"max_per_img" in lib.mehhua_last_cuda_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from aod_meh_hua_b200.scoring import Scorer, pool_topk
    with pytest.raises(_lib.MehhuaError):
        Scorer(get_spec("tiny_retina_coco"))
    with pytest.raises(_lib.MehhuaError):
        pool_topk(torch.rand(10), 3)
    lib = _lib.load()
    out = (C.c_uint32 * 8)()
    rc = lib.mehhua_debug_philox((C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0), out)
    assert rc == _lib.E_NODEVICE
    # the mutual-information baselines: CPU tensors are refused, and the entry point itself reports the missing device
    from aod_meh_hua_b200.mi_baselines import ComputeMI, mutual_information
    members = [[torch.zeros(1, 40, 3, 3)] for _ in range(3)]
    with pytest.raises(_lib.MehhuaError):
        mutual_information(members, 20)
    with pytest.raises(_lib.MehhuaError):
        ComputeMI(*members, nCls=20)
    lv = _lib.LevelArray()
    lv[0].H, lv[0].W, lv[0].A = 3, 3, 2
    ptrs = (C.c_void_p * 3)(8, 8, 8)
    assert lib.mehhua_mi_workspace_bytes(lv, 1, 1) >= 4
    assert lib.mehhua_mi_score_batch(ptrs, 3, lv, 1, 20, 1, None, C.c_void_p(8), C.c_void_p(8), 256, None) == _lib.E_NODEVICE
    # K4's workspace sizes: one block for small pools, the grid-wide form's buffers for large ones
    assert lib.mehhua_pool_topk_workspace_bytes(1000) == 256 and lib.mehhua_pool_topk_workspace_bytes_k(1000, 10) == 256
    big, big_k = lib.mehhua_pool_topk_workspace_bytes(1_000_000), lib.mehhua_pool_topk_workspace_bytes_k(1_000_000, 25_000)
    assert 256 < big_k < big and big >= 2 * 8 * 1_000_000


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "aod_meh_hua_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "import_module(\"oracle" not in src and "__import__(\"oracle" not in src, fn


def test_shard_range_partitions_pool():
    for n in (0, 1, 7, 100000, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for (a, b), (c, d) in zip(cuts[:-1], cuts[1:]):
                assert b == c and 0 <= (b - a) - (d - c) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from aod_meh_hua_b200.pool import gather_scores, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
for n in (11, 64, 1001):
    full = torch.arange(n, dtype=torch.float32) * 0.5 + 1.0
    a, b = shard_range(n, rank, world)
    got = gather_scores(full[a:b].clone(), n, rank, world)
    assert got.shape == (n,) and torch.equal(got, full), (rank, n)
# world-size independence of the selected set: the top-k of the gathered scores is the same on
# every rank and equals the single-process answer
k = 7
top = torch.topk(got, k).indices.sort().values
ref = torch.topk(full, k).indices.sort().values
assert torch.equal(top, ref)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_scores_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert "ok" in out


def test_parse_agg_spec_tokens():
    assert parse_agg_spec("objectSum_scaleMax_classSum") == (0, 2, 0)
    assert parse_agg_spec("classAvg_objectMax_scaleSum") == (2, 0, 1)
    with pytest.raises(KeyError):
        parse_agg_spec("objectMin_scaleMax_classSum")


def test_cycle_files_round_trip(tmp_path):
    """X_L_k / X_U_k / Unc_k .npy (tools/train_RetinaNet.py:249-251) and ResumeCycle
    (utils/functions.py:478-490)."""
    import types
    from aod_meh_hua_b200.pool import ResumeCycle, ResumeCycle_WorkDir, save_cycle
    X_L, X_U = np.array([1, 5, 9]), np.array([0, 2, 3])
    unc = [torch.tensor(float(i)) for i in range(10)]           # the list of 0-d tensors the scorer returns
    save_cycle(str(tmp_path), 3, X_L, X_U, unc)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["Unc_4.npy", "X_L_4.npy", "X_U_4.npy"]
    u = np.load(tmp_path / "Unc_4.npy")
    assert u.dtype == np.float32 and u.shape == (10,)
    cfg = types.SimpleNamespace(work_dir=str(tmp_path))
    assert ResumeCycle(cfg, 3, 4) == (False, False)
    rl, ru = ResumeCycle(cfg, 4, 4)
    assert np.array_equal(rl, X_L) and np.array_equal(ru, X_U) and rl.dtype == X_L.dtype
    rl2, _ = ResumeCycle_WorkDir(str(tmp_path), 6, 4)
    assert np.array_equal(rl2, X_L)


def test_oracle_max_conf_matches_reference():
    import os
    from aod_meh_hua_b200.specs import get_spec
    from aod_meh_hua_b200.synth import SyntheticPool
    from oracle import meh_hua_oracle as O
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kats.npz"))
    for name in ("tiny_retina_coco", "tiny_ssd_voc"):
        spec = get_spec(name)
        batch = SyntheticPool(spec, seed0=20).batch([0, 1, 2])
        per_img, per_lvl = O.get_max_conf(batch["cls_scores"], spec.c_out)
        assert np.array_equal(per_lvl.numpy(), g[f"maxconf_levels_{name}"])
        assert np.array_equal(np.asarray(per_img, dtype=np.float64), g[f"maxconf_{name}"])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path: its own AST-loaded functions where
    /root/reference is mounted, else the oracle port) needs no GPU and prints one JSON line with the
    keys the driver reads."""
    import json
    from oracle import ref_loader as RL
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "tiny_retina_coco", "--steps", "2", "--warmup", "1", "--no-baseline-of-record"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["gpu_launches"] == 0
    assert line["config"]["workload"] == "tiny_retina_coco"
    assert line["cpu_baseline"]["kind"] == ("reference" if RL.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_ctypes_structs_follow_the_header_field_for_field():
    """The ctypes mirrors in _lib.py must list the fields of mehhua_level_t / mehhua_config_t /
    mehhua_buffers_t in the header's order (an ABI drift would silently shift every pointer)."""
    hdr = open(os.path.join(ROOT, "include", "mehhua.h")).read()

    def fields(struct_name):
        body = re.search(r"typedef struct " + struct_name + r"\s*\{(.*?)\}\s*\w+;", hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):          # `int32_t H, W, A;` declares three fields
                m = re.search(r"(\w+)\s*(\[\d+\])?$", part.strip())
                names.append(m.group(1))
        return names

    assert fields("mehhua_buffers") == list(_lib.BUFFER_FIELDS)
    assert fields("mehhua_config") == [f[0] for f in _lib.Config._fields_]
    # the level struct calls the MEH map `lambda` in C and `lam` in Python (keyword)
    assert [n if n != "lambda" else "lam" for n in fields("mehhua_level")] == [f[0] for f in _lib.Level._fields_]
    assert C.sizeof(_lib.Buffers) == 8 * len(_lib.BUFFER_FIELDS)
    ver = int(re.search(r"#define MEHHUA_ABI_VERSION (\d+)", hdr).group(1))
    assert ver == _lib.ABI_VERSION


def test_head_variants_cover_the_reference_scoring_heads():
    """HEAD_VARIANTS: one drop-in class per scoring head of the reference, with the two switches that
    distinguish their scoring methods (thresholds from kwargs, lambda' scaling)."""
    from aod_meh_hua_b200.dropin import HEAD_VARIANTS, B200ScoringMixin, make_variant

    class Base:
        def _get_bboxes(self, *a, **k):
            return "reference"

    want = {"Lambda_L2Net_B200": (False, True), "Lambda_L1Net_B200": (False, True), "Lambda_MSLENet_B200": (False, True),
            "Lambda_L2Net_reverse_B200": (False, True), "Lambda_L2Net_ablation_B200": (True, True),
            "Lambda_L2Net_NoL_B200": (True, False), "Lambda_L2Net_ReLU_B200": (True, False),
            "MyLSSDHead_B200": (False, True), "L_AnchorHead_B200": (False, True)}
    assert {k: v[2:4] for k, v in HEAD_VARIANTS.items()} == want
    assert make_variant("L_AnchorHead_B200", Base).mehhua_params.activation == "relu_plus_one"
    for name in HEAD_VARIANTS:
        cls = make_variant(name, Base)
        assert cls.__name__ == name and issubclass(cls, B200ScoringMixin) and issubclass(cls, Base)
        assert cls.mehhua_thresholds_from_kwargs == want[name][0]
        assert cls.mehhua_params.use_lambda == want[name][1]
        # non-scoring routes fall through to the reference method
        assert cls()._get_bboxes([torch.zeros(1, 9, 2, 2)], [], [], [(8, 8, 3)], [], None, False, False,
                                 isUnc="Epistemic", uPool="Entropy_NoNMS") == "reference"


def test_header_is_plain_c(tmp_path):
    """include/mehhua.h is a C header (C99, no C++ constructs): a C translation unit that includes it
    compiles with gcc -pedantic and links against the shared library."""
    src = tmp_path / "t.c"
    src.write_text('#include "mehhua.h"\n#include <stdio.h>\n'
                   'int main(void) { mehhua_config_t c; mehhua_buffers_t b; (void)c; (void)b;\n'
                   '  printf("%d %d %d\\n", mehhua_abi_version(), (int)sizeof(mehhua_level_t), (int)sizeof(mehhua_buffers_t));\n'
                   '  return mehhua_abi_version() == MEHHUA_ABI_VERSION ? 0 : 1; }\n')
    exe = tmp_path / "t"
    lib_dir = os.path.join(ROOT, "aod_meh_hua_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), "-L", lib_dir, "-lmehhua", f"-Wl,-rpath,{lib_dir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    ver, lvl, bufs = (int(v) for v in run.stdout.split())
    assert ver == _lib.ABI_VERSION and lvl == C.sizeof(_lib.Level) and bufs == C.sizeof(_lib.Buffers)


def test_dropin_composes_with_the_real_reference_class():
    """the mixin in front of the REAL Lambda_L2Net methods (AST-loaded; skipped where
    /root/reference is not mounted).  On CPU tensors no route is taken over, so every call must fall through
    `super()` to the reference's own `_get_bboxes`: the evaluation route returns the reference's det_results
    (checked against the golden minted from the same method), and HEAD_VARIANTS names only classes that exist."""
    from oracle import ref_loader as RL
    if not RL.available():
        pytest.skip("reference tree not mounted")
    import types
    from aod_meh_hua_b200 import dropin
    from aod_meh_hua_b200.specs import get_spec
    from aod_meh_hua_b200.synth import SyntheticPool
    for name, (module, cls, *_r) in dropin.HEAD_VARIANTS.items():
        src = open(os.path.join(RL.REF_ROOT, "mmdet/models/dense_heads", module + ".py")).read()
        assert re.search(rf"^class {cls}\(", src, flags=re.M), (name, module, cls)
    spec = get_spec("tiny_retina_voc")
    ns = RL.base_namespace()
    fns = RL.load_methods("mmdet/models/dense_heads/Lambda_L2.py", "Lambda_L2Net",
                          ["_get_bboxes", "ComputeObjUnc", "AggregateObjScaleUnc", "ComputeScaleUnc", "AggregateScaleUnc"], ns)
    Ref = type("Lambda_L2Net", (object,), dict(fns))
    Head = dropin.make_variant("Lambda_L2Net_B200", Ref)
    assert Head.__mro__[1] is dropin.B200ScoringMixin and Head.__mro__[2] is Ref
    h = Head()
    h.cls_out_channels, h.last_activation = spec.c_out, "relu"
    h.test_cfg = RL._Cfg(nms_pre=spec.nms_pre, min_bbox_size=0, score_thr=spec.score_thr,
                         nms=dict(type="nms", iou_threshold=spec.nms_iou), max_per_img=spec.max_per_img)
    d2b = ns["delta2bbox"]
    h.bbox_coder = types.SimpleNamespace(means=(0., 0., 0., 0.), stds=spec.target_stds,
                                         decode=lambda rois, deltas, max_shape=None: d2b(rois, deltas, (0., 0., 0., 0.), tuple(spec.target_stds), max_shape))
    batch = SyntheticPool(spec, seed0=20).batch([0, 1])
    sf = [np.asarray(v, dtype=np.float32) for v in batch["scale_factors"]]
    dets = h._get_bboxes(batch["cls_scores"], batch["bbox_preds"], batch["anchors"], batch["img_shapes"], sf, None, True, True,
                         isUnc=None, isEval=True)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "retina_voc.npz"))
    for b, (d, l) in enumerate(dets):
        assert np.array_equal(d.numpy(), g[f"dets_{b}"]) and np.array_equal(l.numpy(), g[f"labels_{b}"])
