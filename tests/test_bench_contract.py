"""bench.py's contract that can be checked without a GPU: the reference arm (the CPU restatement timed on the host cores)
prints ONE JSON line with the keys the driver reads, and the kernel names the roofline reports follow the dispatcher."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fse_roundtrip_GBps_uncompressed" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["dtype"] == "u8" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("c1")


def test_kernel_names_follow_the_dispatcher():
    sys.path.insert(0, ROOT)
    import bench
    k = bench.kernel_names(0, 0, 0, 65536)                  # c4 on one GPU: wide 32-bit entries
    assert (k["encode"], k["decode"]) == ("k_encode128_blocks", "k_decode128_blocks")
    assert bench.kernel_names(0, 0, 0, 4096)["decode"] == "k_decode128c_blocks"        # c2: too few blocks per warp slot
    assert bench.kernel_names(0, 0, 12, 16384)["decode"] == "k_decode128c_blocks"      # table_log 12: compact keeps more warps
    assert bench.kernel_names(0, 0, 9, 16384)["decode"] == "k_decode128_blocks"        # small tables: as many warps either way
    assert bench.kernel_names(1, 0, 11, 65536)["encode"] == "k_encode_sh_global"
    assert bench.kernel_names(0, 8192, 0, 65536)["decode"] == "k_decode_sh_blocks"
