"""CPU tests of the measurement tooling: the ncu launch-list parser behind `roofline.traffic` and bench.py's byte model."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CSV = '''==PROF== Connected to process 1
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","1","python","h","void ia_bwd_kernel<__nv_bfloat16, 1, 2, 3>(const V *, int)","1","7","(256, 1, 1)","(5120, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Mbyte","90.5"
"0","1","python","h","void ia_bwd_kernel<__nv_bfloat16, 1, 2, 3>(const V *, int)","1","7","(256, 1, 1)","(5120, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","Mbyte","9.5"
"0","1","python","h","void ia_bwd_kernel<__nv_bfloat16, 1, 2, 3>(const V *, int)","1","7","(256, 1, 1)","(5120, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","56.0"
"1","1","python","h","ia_param_grad_kernel(const float *, int)","1","7","(128, 1, 1)","(1, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Kbyte","12"
"1","1","python","h","ia_param_grad_kernel(const float *, int)","1","7","(128, 1, 1)","(1, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","byte","0"
"1","1","python","h","ia_param_grad_kernel(const float *, int)","1","7","(128, 1, 1)","(1, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","5400"
"2","1","python","h","rt_finalize_kernel(const float *, int)","1","7","(128, 1, 1)","(32, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","3.0"
"3","1","python","h","rt_finalize_kernel(const float *, int)","1","7","(128, 1, 1)","(32, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","3.0"
'''


def test_ncu_traffic_groups_units_and_steps(tmp_path):
    src = os.path.join(tmp_path, "launches.csv")
    with open(src, "w") as f:
        f.write(CSV)
    oj, ot = os.path.join(tmp_path, "t.json"), os.path.join(tmp_path, "t.txt")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), src, oj, ot], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.load(open(oj))
    assert d["steps_seen"] == 2                                  # two launches of the once-per-step kernel
    g = d["groups"]["in_act_bwd"]                                # ia_bwd + ia_param_grad belong to one group
    assert g["launches_per_step"] == 1.0
    assert abs(g["us_per_step"] - (56.0 + 5.4) / 2) < 1e-6
    assert g["dram_bytes_per_step"] == int((100e6 + 12e3) / 2)
    assert g["dram_bytes_per_launch"] == int((100e6 + 12e3) / 2)
    assert d["groups"]["recon_tail_fwd"]["launches_per_step"] == 1.0
    assert "in_act_bwd" in open(ot).read()


def test_bench_byte_model_follows_the_io_width():
    sys.path.insert(0, ROOT)
    import bench
    f32 = bench.alg_bytes("eb4", 32, 380, 16, 4)
    b16 = bench.alg_bytes("eb4", 32, 380, 16, 2)
    assert b16["in_act_fwd"] * 2 == f32["in_act_fwd"] and b16["in_act_bwd"] * 2 == f32["in_act_bwd"]
    assert b16["recon_tail_fwd"] == f32["recon_tail_fwd"]       # the tail stays fp32
    # G2 = 24.1 / 36.2 MB per sample forward / backward in fp32, recon tail 3.91 / 2.62 MB
    # (G2 = InstanceNorm epilogues + the final tanh)
    assert abs((f32["in_act_fwd"] + f32["tanh_fwd"]) / 32 / 1e6 - 24.1) < 0.1
    assert abs((f32["in_act_bwd"] + f32["tanh_bwd"]) / 32 / 1e6 - 36.2) < 0.1
    assert abs(f32["recon_tail_fwd"] / 32 / 1e6 - 3.91) < 0.01
    committed = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(committed):
        assert "in_act_bwd" in bench.TRAFFIC and bench.TRAFFIC["in_act_bwd"] > 0
