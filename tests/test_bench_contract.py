"""CPU tests of bench.py's bookkeeping: the algorithmic-byte model equals SURVEY.md 8(d) / BASELINE.md section 2, and the argument parser exposes
the contract's flags."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_the_survey_totals():
    sys.path.insert(0, ROOT)
    import bench
    # 2^16/None 280.3 MB, 2^20/None 4484 MB, 2^20/Quadratic 4906 MB
    assert round(bench.algorithmic_bytes(16, 1)["_total_survey"] / 1e6, 1) == 280.3
    assert round(bench.algorithmic_bytes(20, 1)["_total_survey"] / 1e6) == 4484
    ab = bench.algorithmic_bytes(20, 2)
    assert round(ab["_total_survey"] / 1e6) == 4906
    # per-stage figures of BASELINE.md (MB): interpolate 117, LDE 528, constraints 151, composition iNTT 67, composition LDE 151, OOD 76
    assert [round(ab[k] / 1e6) for k in ("ntt.interpolate_trace", "ntt.lde_trace", "constraints", "ntt.interpolate_comp", "ntt.lde_comp")] == [117, 528, 151, 67, 151]
    assert round(ab["ood"] / 1e6) in (75, 76)
    # the kernel families of this backend partition the same traffic up to the fused leaf/tree split
    fam = sum(v for k, v in ab.items() if not k.startswith("_") and k != "combine")
    assert abs(fam - ab["_total_survey"]) / ab["_total_survey"] < 0.08


def test_cli_flags_of_the_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload"):
        assert flag in out
    for w in ("latency", "batch", "wide", "verify", "air"):
        assert w in out


def test_ntt_gbps_metric():
    sys.path.insert(0, ROOT)
    import bench
    acc = {"ntt.lde_trace": [1.1409, 2], "ntt.interpolate_trace": [0.1595, 2], "deep": [0.63, 1]}
    r = bench.ntt_gbps_from(acc, 20, 2)
    assert set(r) == {"ntt.lde_trace", "ntt.interpolate_trace"}
    assert r["ntt.lde_trace"] == {"transforms": 56, "gbps": round(16 * 56 * (1 << 20) / 1.1409 / 1e6, 1)}
    assert 700 < r["ntt.interpolate_trace"]["gbps"] < 760
    assert bench.ntt_gbps_from({}, 16, 1) == {}
