"""CPU tests of the measurement harness (bench.py / oracle/cpu_baseline.py): the JSON contract of the reference arm, the
bounded CPU-arm fallback chain, the algorithmic-FLOP formula the roofline uses, and the synthetic batch layout."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from oracle import cpu_baseline  # noqa: E402


def test_per_layer_extrapolation_floor():
    d, note = cpu_baseline._per_layer(12.5, 19.9)           # a clean pair measured on the GPU box's host
    assert abs(d - 7.4) < 1e-9 and note == ""
    d, note = cpu_baseline._per_layer(55.7, 44.5)           # host noise: the 2-layer step came out faster
    assert abs(d - 0.25 * 44.5) < 1e-9 and "floor" in note
    d, note = cpu_baseline._per_layer(10.0, 10.5)           # difference below a quarter of t2
    assert d == 0.25 * 10.5 and note


def test_train_flops_formula_matches_survey():
    """12.71 TFLOP per CogACT-7B sample (S = 309); pi0 12.1; the OFT variants add their heads."""
    w = bench.WORKLOADS["cogact_7b"]
    assert abs(bench.train_flops_per_sample(w, 309) / 1e12 - 12.708) < 5e-3
    assert abs(bench.train_flops_per_sample(bench.WORKLOADS["pi0_2b"], 867) / 1e12 - 12.109) < 5e-3
    assert bench.train_flops_per_sample(bench.WORKLOADS["oft_discrete_7b"], 365) > bench.train_flops_per_sample(w, 365)
    # training = 3x forward and the decoder dominates: 6 * params * tokens within 10 % for the 7B decoder
    L = w["llm"]
    dec_params = L["num_hidden_layers"] * (2 * L["hidden_size"] * L["hidden_size"] * (1 + 4 / 28) +
                                           3 * L["hidden_size"] * L["intermediate_size"])
    assert 0.85 < 6 * dec_params * 309 / bench.train_flops_per_sample(w, 309) < 1.0


@pytest.mark.parametrize("name", ["cogact_7b", "oft_discrete_7b", "pi0_2b", "memvla_7b"])
def test_make_batch_layout(name):
    w = dict(bench.WORKLOADS[name], batch=3)
    b = bench.make_batch(w, rank=1, pinned=False)
    b2 = bench.make_batch(w, rank=1, pinned=False)
    for k, v in b.items():
        if hasattr(v, "shape"):
            assert v.shape[0] == 3 and (v == b2[k]).all(), k           # seeded: the same batch whoever builds it
    if w.get("kind") == "pi0":
        assert b["images"].shape == (3, w["n_cam"], 3, 224, 224) and b["actions"].shape == (3, 50, 32)
    else:
        assert (b["input_ids"][:, 1] == -200).all() and b["images"].shape[-3:] == (3, 224, 224)
    if w.get("kind") == "oft_discrete":
        lab = b["labels"]
        assert ((lab >= 0).sum(1) == w["extra_tokens"]).all()
        assert int(lab.max()) < w["llm"]["vocab_size"] and int(lab[lab >= 0].min()) >= w["llm"]["vocab_size"] - w["num_bins"]
    if w.get("kind") == "memvla":
        assert len(b["indexes"]) == 3


def test_cpu_arm_fallback_chain(monkeypatch):
    """reference kind -> port batch 4 (-> port batch 1 when the reference is not preferred); every attempt is a child
    process under a limit; the first JSON line wins; nothing finishing yields value None, never an exception."""
    calls = []

    class R:
        def __init__(self, rc, out):
            self.returncode, self.stdout, self.stderr = rc, out, "boom"

    def fake_run(cmd, **kw):
        calls.append((cmd[-1], cmd[-2], kw["timeout"], kw["env"].get("OMP_NUM_THREADS"), kw["env"]["CUDA_VISIBLE_DEVICES"]))
        kind = cmd[-1]
        if kind == "reference":
            raise subprocess.TimeoutExpired(cmd, kw["timeout"])
        return R(0, 'noise\n{"value": 0.02, "kind": "port", "sample": "x"}\n')

    monkeypatch.setenv("OMP_NUM_THREADS", "1")            # what torchrun exports: must not reach the child
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    out = bench.cpu_baseline({}, 309, workload="cogact_7b", limit_s=100.0, prefer_reference=True)
    assert out["value"] == 0.02 and [c[0] for c in calls] == ["reference", "port"]
    assert calls[0][2] == 225.0 and calls[1][2] == 100.0 and all(c[3] is None and c[4] == "" for c in calls)

    calls.clear()
    monkeypatch.setattr(bench.subprocess, "run", lambda cmd, **kw: (calls.append(cmd[-2]), R(1, ""))[1])
    out = bench.cpu_baseline({}, 309, workload="cogact_7b", limit_s=10.0)
    assert out["value"] is None and "unavailable" in out["sample"] and calls == ["4", "1"]


def test_reference_arm_json_contract_tiny():
    """`bench.py --impl reference` end to end on the tiny workload (CPU only): one JSON line with the keys the driver
    reads; under torchrun ranks other than 0 print nothing."""
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "cogact_tiny", "--steps", "1",
           "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=120, cwd=str(ROOT), env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
