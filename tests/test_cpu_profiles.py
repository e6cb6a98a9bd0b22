"""CPU: the committed measurement records stay derivable from the committed raw profiler exports (nothing typed by hand)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda *a: os.path.join(ROOT, "profiles", *a)


def test_ncu_traffic_record_is_reproduced_from_the_raw_exports(tmp_path):
    out = tmp_path / "traffic.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_ncu_traffic.py"),
                    P("r02_ncu_full_gemm_shapes_final_c28.csv"), P("r02_ncu_full_attention_bwd_final_c28.csv"),
                    P("r02_gemm_shapes_final_c28.json"), str(out), P("r02_ncu_full_attention_b8_final_c29.csv")],
                   check=True, capture_output=True, cwd=ROOT)
    new, old = json.load(open(out)), json.load(open(P("r02_ncu_traffic.json")))
    assert new["gemm"]["dram_bytes_per_launch"] == old["gemm"]["dram_bytes_per_launch"]
    assert new["gemm"]["algorithmic_bytes_per_launch"] == old["gemm"]["algorithmic_bytes_per_launch"]
    assert [r["name"] for r in new["gemm_launch_table"]] == [r["name"] for r in old["gemm_launch_table"]]
    for a, b in zip(new["gemm_launch_table"], old["gemm_launch_table"]):
        assert a["dram_read_bytes"] == b["dram_read_bytes"] and a["ms_ncu"] == b["ms_ncu"]
        # written traffic can never be below ~the output size; read traffic never below the operands the launch must fetch once
        assert a["dram_write_bytes"] > 0 and a["traffic_over_algorithmic"] > 0.9
    assert new["attention"]["dram_bytes_per_launch"] == old["attention"]["dram_bytes_per_launch"]
    assert "split" in new["attention"]["kernel"]          # the bench-shape capture is the bounded-score kernel


def test_bench_reads_the_round2_traffic_record():
    sys.path.insert(0, ROOT)
    import bench
    t = bench._ncu_traffic()
    assert t and t.get("file", "").endswith("r02_ncu_traffic.json")
    assert t["gemm"]["kernel"] == "gemm_bf16_2cta_kernel" and t["gemm"]["dram_bytes_per_launch"] > t["gemm"]["algorithmic_bytes_per_launch"]
