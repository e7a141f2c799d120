"""bench.py's roofline bookkeeping, without a GPU: the ncu capture recorded in profiles/traffic.json is only used when it
was taken on the kernel sources in the tree (a stale capture must print null, not a number), and the issue-port model
(2 issue cycles per FP64 warp instruction + 1 per other) adds up."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_committed_capture_is_of_the_committed_kernels():
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    if d["kernel_sources_sha"] != bench.kernel_sources_sha():      # kernels edited since the capture: bench.py must say so
        assert bench.ncu_traffic("v_pass_stream")[0] is None and bench.ncu_issue_cycles(["v_pass_stream"]) is None
        return
    for k in ("v_pass_stream", "h_pass_epilogue"):
        traffic, src = bench.ncu_traffic(k)
        assert traffic and traffic > 124416000 / 2, (k, traffic, src)     # at least the image itself
        assert d[k + "_warp_instr_fp64"] < d[k + "_warp_instr"]


def test_issue_port_cycles_follow_the_model():
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    if d["kernel_sources_sha"] != bench.kernel_sources_sha():
        return
    want = sum((d[k + "_warp_instr"] + d[k + "_warp_instr_fp64"]) / (4.0 * 148) for k in ("v_pass_stream", "h_pass_epilogue"))
    got = bench.ncu_issue_cycles(["v_pass_stream", "h_pass_epilogue"])
    assert got is not None and abs(got - want) < 1e-6
    # the cfg2 taps alone: 895 263 840 DMUL+DADD pairs per image = 2 * pairs / 32 warp instructions, each 2 issue cycles
    taps_cycles = 2 * 895263840 / 32 * 2 / (4.0 * 148)
    assert taps_cycles < got < 2.0 * taps_cycles


def test_stale_capture_is_refused(tmp_path, monkeypatch):
    stale = dict(json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))), kernel_sources_sha="0" * 16)
    (tmp_path / "profiles").mkdir()
    json.dump(stale, open(tmp_path / "profiles" / "traffic.json", "w"))
    monkeypatch.setattr(bench, "kernel_sources_sha", lambda: "f" * 16)
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    traffic, why = bench.ncu_traffic("v_pass_stream")
    assert traffic is None and "other kernel sources" in why
    assert bench.ncu_issue_cycles(["v_pass_stream"]) is None
