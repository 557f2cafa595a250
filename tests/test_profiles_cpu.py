"""The committed ncu evidence must belong to the committed kernels: bench.py reports `roofline.traffic` only while the kernel source still
hashes to what `profiles/ncu_traffic.json` names (otherwise it prints null) — this test makes a stale entry visible on every CPU run."""
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_traffic_entries_name_the_committed_kernel_source_and_profile():
    rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
    assert {"mlp_tc_bwd", "mlp_tc_fwd"} <= set(rec)
    for name, r in rec.items():
        src = os.path.join(ROOT, r["source"])
        assert os.path.exists(src), (name, r["source"])
        assert hashlib.sha256(open(src, "rb").read()).hexdigest() == r["source_sha256"], \
            f"{name}: {r['source']} changed after the ncu capture in {r['profile']} — recapture (B200_PROFILING recipe) or mark the entry stale"
        assert os.path.exists(os.path.join(ROOT, r["profile"])), (name, r["profile"])
        # DRAM bytes of a launch can undercut the algorithmic bytes (L2 keeps part of the output) but never exceed them by much:
        # traffic well above algorithmic would mean wasted re-reads
        assert 0.5 * r["algorithmic_bytes"] <= r["dram_bytes"] <= 1.1 * r["algorithmic_bytes"], (name, r["dram_bytes"], r["algorithmic_bytes"])
