"""Host-side checks of the drop-in overlay that need no GPU: import surface, state_dict keys,
initial-weight parity with the reference (same seed, same construction order), C-ABI symbols."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OVERLAY = os.path.join(ROOT, "pfotgnrec_b200", "overlay")


@pytest.fixture(scope="module")
def overlay():
    sys.path.insert(0, OVERLAY)
    for m in [k for k in sys.modules if k.split(".")[0] in ("model", "modules", "utils", "evaluation")]:
        del sys.modules[m]
    import model.tgn as tgn_mod
    import utils.utils as utils_mod
    yield tgn_mod, utils_mod
    sys.path.remove(OVERLAY)


def _build(tgn_mod, z, device="cpu"):
    torch.manual_seed(11)
    return tgn_mod.TGN(neighbor_finder=None, node_features=z["node_feat"], edge_features=z["st_edge_features"].copy(),
                       device=torch.device(device), n_layers=int(z["cfg_n_layers"]), n_heads=2, dropout=0.0,
                       use_memory=bool(z["cfg_use_memory"]), message_dimension=100, memory_dimension=int(z["cfg_d"]),
                       memory_update_at_start=True, embedding_module_type=str(z["cfg_embedding"]),
                       message_function=str(z["cfg_msg_fn"]) if "cfg_msg_fn" in z else "identity",
                       aggregator_type=str(z["cfg_aggregator"]) if "cfg_aggregator" in z else "last",
                       memory_updater_type=str(z["cfg_updater"]), n_neighbors=int(z["cfg_n_neighbors"]),
                       mean_time_shift_src=z["cfg_shift"][0], std_time_shift_src=z["cfg_shift"][1],
                       mean_time_shift_dst=z["cfg_shift"][2], std_time_shift_dst=z["cfg_shift"][3],
                       use_destination_embedding_in_message=bool(z["cfg_dst_emb"]),
                       use_source_embedding_in_message=bool(z["cfg_src_emb"]) if "cfg_src_emb" in z else False,
                       dyrep=bool(z["cfg_dyrep"]))


@pytest.mark.parametrize("tag", ["ours", "jodie", "dyrep", "tgat2", "mlp_mean", "srcemb", "gsum", "gsum2"])
def test_initial_weights_and_keys_match_reference(overlay, tag):
    tgn_mod, _ = overlay
    z = load_golden(f"tgn_{tag}.npz")
    tgn = _build(tgn_mod, z)
    sd = tgn.state_dict()
    ref_keys = {k[2:] for k in z if k.startswith("w_")}
    mine = {k for k in sd if not re.match(r"(memory\.|memory_updater\.memory\.|embedding_module\.memory\.)", k)}
    assert mine == ref_keys
    for k in ref_keys:
        assert np.array_equal(sd[k].numpy(), z["w_" + k]), k
    if bool(z["cfg_use_memory"]):
        for alias in ("memory.memory", "memory_updater.memory.memory", "embedding_module.memory.last_update"):
            assert alias in sd


def test_import_surface(overlay):
    tgn_mod, utils_mod = overlay
    for name in ("EarlyStopMonitor", "RandEdgeSampler", "get_neighbor_finder", "MergeLayer", "MLP", "NeighborFinder"):
        assert hasattr(utils_mod, name)
    for name in ("compute_temporal_embeddings_p", "compute_temporal_embeddings", "set_neighbor_finder"):
        assert hasattr(tgn_mod.TGN, name)
    import evaluation as ev_mod            # overlay/evaluation.py shadows the reference's top-level module
    import inspect
    assert list(inspect.signature(ev_mod.eval_recommendation).parameters) == [
        "tgn", "data", "full_data", "batch_size", "n_neighbors", "upper_u", "period", "is_test_run", "EVAL"]
    import modules.memory as mm
    for name in ("__init_memory__", "detach_memory", "backup_memory", "restore_memory", "get_memory",
                 "set_memory", "get_last_update", "store_raw_messages", "clear_messages"):
        assert hasattr(mm.Memory, name)


def test_c_abi_exports_every_declared_symbol():
    from pfotgnrec_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pfo_b200.h")).read()
    declared = set(re.findall(r"\b(pfo_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.pfo_abi_version() == _lib.ABI_VERSION == 3


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pfotgnrec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)


def _reference_arm(extra=()):
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--users", "300", "--items", "40", "--events", "3000", "--days", "20",
                          "--bs", "64", "--ref-budget", "6", "--eval-budget", "2", *extra],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def _check_reference_line(d, kind):
    assert d["impl"] == "reference" and d["metric"] == "train_events_per_sec" and d["unit"] == "events/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["global_batch"] == 64


def test_bench_reference_arm_json_contract_port():
    """`bench.py --impl reference --ref-kind port` (the oracle port on the host cores, the fallback when no reference
    tree is staged) prints ONE JSON line with the contract keys the driver reads."""
    _check_reference_line(_reference_arm(["--ref-kind", "port"]), "port")


def test_bench_reference_arm_runs_the_unmodified_reference():
    """With a reference tree present (/root/reference here, baseline/_ref on the GPU box) the arm drives the reference's
    own modules and the text of main.py:167-394 (kind "reference"), training and evaluation."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from stage_reference import ref_root
    if ref_root() is None:
        pytest.skip("no reference tree staged")
    d = _reference_arm()
    _check_reference_line(d, "reference")
    ev = d["eval"]
    assert ev["metric"] == "eval_users_per_sec" and ev["value"] > 0 and ev["cpu_baseline"]["kind"] == "reference"


def test_staged_reference_is_byte_identical():
    """baseline/_ref (git-ignored, shipped to the GPU box) is a byte-for-byte copy of the reference's python files."""
    import hashlib
    import json
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not (os.path.isdir("/root/reference") and os.path.exists(os.path.join(ref, "MANIFEST.json"))):
        pytest.skip("needs both /root/reference and the staged copy")
    for rel, sha in json.load(open(os.path.join(ref, "MANIFEST.json"))).items():
        assert hashlib.sha256(open(os.path.join("/root/reference", rel), "rb").read()).hexdigest() == sha
        assert hashlib.sha256(open(os.path.join(ref, rel), "rb").read()).hexdigest() == sha


def test_reference_on_disk_format_round_trip(tmp_path):
    """synth.write_reference_format -> synth.read_reference_format reproduces the stream: ids, times, edge features,
    day rows, portfolio CSR and both price tables (the files are the ones reference utils/data.py:20-25, main.py:88-89
    and evaluation.py:41-43 load)."""
    from pfotgnrec_b200.synth import make_stream, write_reference_format, read_reference_format
    st = make_stream(n_users=120, n_items=25, n_events=900, n_days=9, seed=3, ts_mode="nbg")
    write_reference_format(st, str(tmp_path), period="30")
    rt = read_reference_format(str(tmp_path), period="30")
    assert rt.n_users == st.sources.max() and rt.n_items == st.n_items and rt.codes == st.codes
    for k in ("sources", "destinations", "timestamps", "edge_idxs", "edge_features", "port_ptr", "port_items",
              "prices_future", "prices_past"):
        assert np.array_equal(getattr(rt, k), getattr(st, k)), k
    used = np.unique(st.day_idx)                  # day rows: same price rows for every interaction
    assert [rt.day_keys[i] for i in rt.day_idx[:50]] == [st.day_keys[i] for i in st.day_idx[:50]]
    assert np.array_equal(rt.prices_future[rt.day_idx], st.prices_future[st.day_idx]) and used.size > 0


def test_bench_stream_cursor_wraps_inside_the_region():
    """bench.py's batch positions: consecutive, inside [lo, hi), wrapping to lo instead of running past the end (an
    8-GPU run with 65 536-event global batches needs more events than the stream holds after the start offset)."""
    sys.path.insert(0, ROOT)
    from bench import StreamCursor
    c = StreamCursor(2_000_000, 5_000_000)
    seen = [c.take(65536) for _ in range(120)]
    assert all(2_000_000 <= s and e <= 5_000_000 and e - s == 65536 for s, e in seen)
    assert seen[0] == (2_000_000, 2_065_536) and seen[1][0] == seen[0][1]
    wraps = [i for i in range(1, 120) if seen[i][0] != seen[i - 1][1]]
    assert wraps == [45, 90] and all(seen[i][0] == 2_000_000 for i in wraps)       # (5M - 2M) // 65536 = 45 batches per lap
    with pytest.raises(SystemExit):
        StreamCursor(0, 100).take(101)


@pytest.mark.parametrize("tag", ["ours", "tgn", "jodie", "dyrep", "tgat2", "mlp_mean", "srcemb", "gsum", "gsum2"])
def test_engine_binds_every_parameter_of_every_config(overlay, tag):
    """Host-side plumbing between the drop-in TGN and the step engine, no kernel launched: the engine is built from
    the overlay model's configuration and every parameter it packs for the kernels exists under the reference's name
    (a renamed container attribute would only surface on the GPU otherwise)."""
    tgn_mod, _ = overlay
    z = load_golden(f"tgn_{tag}.npz")
    tgn = _build(tgn_mod, z)
    eng = tgn._get_engine()
    params = tgn._params()
    names = eng.param_names()
    assert all(n in params for n in names), [n for n in names if n not in params]
    flat = eng._pack(params)
    assert len(flat) == len(names) and all(t.dtype == torch.float32 for t in flat)
    assert eng.cfg.embedding == str(z["cfg_embedding"]) and eng.cfg.use_memory == bool(z["cfg_use_memory"])
    if "cfg_msg_fn" in z:
        assert eng.cfg.message_fn == (str(z["cfg_msg_fn"]) if eng.cfg.use_memory else "identity")
        assert eng.cfg.aggregator == str(z["cfg_aggregator"])
    trainable = {k for k, p in tgn.named_parameters() if p.requires_grad}
    unused = {"memory_updater.layer_norm.weight", "memory_updater.layer_norm.bias"}       # never applied (reference too)
    assert trainable - unused <= set(names) | {n.replace(".mlp.", ".layers.") for n in names}


def _state_dict_case():
    z = load_golden("state_dict.npz")
    sd = {str(k): torch.tensor(z["sd_" + str(k)]) for k in z["sd_keys"]}
    return z, sd


def build_for_state_dict(tgn_mod, z, device, nf=None, seed=99):
    d, n = int(z["cfg_d"]), int(z["cfg_n"])
    torch.manual_seed(seed)
    return tgn_mod.TGN(neighbor_finder=nf, node_features=z["node_feat"], edge_features=z["st_edge_features"].copy(),
                       device=torch.device(device), n_layers=1, n_heads=2, dropout=0.0, use_memory=True,
                       message_dimension=100, memory_dimension=d, memory_update_at_start=True,
                       embedding_module_type="graph_attention", message_function="identity", aggregator_type="last",
                       memory_updater_type="gru", n_neighbors=n, mean_time_shift_src=0.0, std_time_shift_src=1.0,
                       mean_time_shift_dst=0.0, std_time_shift_dst=1.0, use_destination_embedding_in_message=False,
                       use_source_embedding_in_message=False, dyrep=False)


def test_overlay_loads_a_reference_checkpoint(overlay):
    """A state_dict written by the UNMODIFIED reference after two Adam steps (tests/golden/state_dict.npz: every key,
    including the three aliases of the memory buffers, the layer_norm the reference never applies and time_encoder under
    two names) loads into the drop-in with strict=True; all tensors arrive, and the aliases still share one storage
    with the engine's dense state."""
    tgn_mod, _ = overlay
    z, sd = _state_dict_case()
    tgn = build_for_state_dict(tgn_mod, z, "cpu")
    assert set(tgn.state_dict().keys()) == set(sd.keys())
    res = tgn.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    got = tgn.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    st = tgn.memory.state
    assert st.memory.data_ptr() == tgn.memory.memory.data_ptr() == tgn.memory_updater.memory.memory.data_ptr() \
        == tgn.embedding_module.memory.memory.data_ptr()
    assert torch.equal(st.memory, sd["memory.memory"]) and torch.equal(st.last_update, sd["memory.last_update"])
    assert float(sd["memory.memory"].abs().sum()) > 0


def test_packed_batch_layout_round_trip():
    """The one-buffer batch layout of the trainer (host staging buffer == static device buffers): fields are 16-byte
    aligned, disjoint, in the same place for any portfolio length, and the views round-trip their columns."""
    import torch
    from pfotgnrec_b200.trainer import PfoTrainer
    B = 37
    lay_a, n_a = PfoTrainer._pack_layout(B, 5)
    lay_b, n_b = PfoTrainer._pack_layout(B, 5 * B + 1)
    assert n_a <= n_b and n_a % 16 == 0
    spans = []
    for k, (off, n, dt) in lay_a.items():
        assert off % 16 == 0 and lay_b[k][0] == off          # offsets depend on B only: a short host buffer is a prefix
        spans.append((off, off + n * torch.empty((), dtype=dt).element_size()))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= n_a
    assert max(lay_a, key=lambda k: lay_a[k][0]) == "port_items"
    buf = torch.zeros(n_a, dtype=torch.uint8)
    v = PfoTrainer._packed_views(buf, B, 5)
    assert v["src"].dtype == torch.int32 and v["ts"].dtype == torch.float64 and v["port_ptr"].shape[0] == B + 1
    v["ts"].copy_(torch.arange(B, dtype=torch.float64) * 1e9 + 0.5)
    v["ev"].copy_(torch.arange(B, dtype=torch.int64) + (1 << 40))
    v["port_items"].copy_(torch.arange(5, dtype=torch.int32))
    big = torch.zeros(n_b, dtype=torch.uint8)
    big[:n_a].copy_(buf)                                      # what train_step_host does with the device buffer
    w = PfoTrainer._packed_views(big, B, 5 * B + 1)
    assert torch.equal(w["ts"], v["ts"]) and torch.equal(w["ev"], v["ev"]) and torch.equal(w["port_items"][:5], v["port_items"])
    assert int(w["src"].abs().sum()) == 0


def test_half_turn_cosine_constants_hold_their_error_bound():
    """The cosine-only form of the neighbour forward kernel (csrc/pfo_math.cuh::pfo_cosf_half), restated on the host with
    the constants parsed out of the header and exactly-rounded fused multiply-adds: |error| < 2e-7 against libm from
    day-scale arguments up to 1e12 rad (TimeEncode arguments on YYYYMMDDhhmmss timestamps reach ~1e10), and the reduced
    argument stays inside the interval the polynomial was fitted on."""
    import re
    from fractions import Fraction as Fr
    src = open(os.path.join(ROOT, "pfotgnrec_b200", "csrc", "pfo_math.cuh")).read()
    body = src[src.index("float pfo_cosf_half(float x)"):]
    body = body[:body.index("\n}\n")]
    inv_pi, magic = [float(v) for v in re.search(r"fma\(xd, ([0-9.eE+-]+), ([0-9.eE+-]+)\)", body).groups()]
    pi_hi, pi_lo = [float(v) for v in re.findall(r"fma\(-kd, ([0-9.eE+-]+), ", body)]
    c0, c1 = [float(v) for v in re.search(r"fmaf\(u, ([0-9.eE+-]+)f, ([0-9.eE+-]+)f\)", body).groups()]
    rest = [float(v) for v in re.findall(r"fmaf\(u, c, ([0-9.eE+-]+)f\)", body)]
    coef = [float(np.float32(v)) for v in [c0, c1] + rest]
    assert len(coef) == 6 and coef[-1] == 1.0 and abs(inv_pi * pi_hi - 1.0) < 1e-15 and abs(pi_hi + pi_lo - np.pi) < 1e-15

    def fma(a, b, c):                                   # one rounding, like the device fma
        return float(Fr(a) * Fr(b) + Fr(c))

    def f32(v):
        return float(np.float32(v))

    rng = np.random.default_rng(0)
    for scale in (1.0, 1e3, 1e7, 1e10, 1e12):
        worst, r_max = 0.0, 0.0
        for x in ((rng.random(1500) * 2 - 1) * scale).astype(np.float32):
            xd = float(x)
            t = fma(xd, inv_pi, magic)
            odd = int(np.float64(t).view(np.int64)) & 1
            kd = t - magic
            r = fma(-kd, pi_lo, fma(-kd, pi_hi, xd))
            u = f32(f32(r) * f32(r))
            c = f32(fma(u, coef[0], coef[1]))
            for ck in coef[2:]:
                c = f32(fma(u, c, ck))
            worst = max(worst, abs((-c if odd else c) - float(np.cos(xd))))
            r_max = max(r_max, abs(r))
        assert worst < 2e-7, (scale, worst)
        assert r_max < (np.pi / 2) * 1.0003, (scale, r_max)
