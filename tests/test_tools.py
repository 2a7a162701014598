"""The evidence pipeline's parsers (tools/ncu_summary.py) on tiny CSV fixtures: profiles/ must not silently go wrong."""
import importlib.util
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _load(name):
    spec = importlib.util.spec_from_file_location(name, ROOT / "tools" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_ncu_summary_parses_raw_page_and_launch_list(tmp_path):
    ns = _load("ncu_summary")
    raw = tmp_path / "raw.csv"
    raw.write_text(
        '"ID","Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread","launch__shared_mem_per_block_dynamic"\n'
        '"","","us","Mbyte","Kbyte","register/thread","Kbyte/block"\n'
        '"0","void ggrt::color_kernel<4, 0>(ggrt::View, const float *)","20.5","99.6","4111.1","72","76.8"\n'
        '"1","ggrt::render_forward_kernel(ggrt::View)","84064","17.6","6.6","40","0"\n')
    k = ns.read_raw(str(raw))
    assert set(k) == {"color_kernel<4, 0>", "render_forward_kernel"}
    c = k["color_kernel<4, 0>"]
    assert abs(c["ncu_time_us"] - 20.5) < 1e-9 and abs(c["dram_read_MB"] - 99.6) < 1e-9
    assert abs(c["dram_write_MB"] - 4.1111) < 1e-6 and c["registers"] == 72
    assert c["dram_traffic_bytes"] == int(round((99.6 + 4.1111) * 1e6))
    launches = tmp_path / "launches.csv"
    launches.write_text(
        "==PROF== Connected to process 1\n"
        '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"\n'
        '"0","1","python","h","void at::fill_kernel<float>(int)","1","7","(128, 1, 1)","(1, 1, 1)","0","10.0","s","gpu__time_duration.sum","ns","38240"\n'
        '"1","1","python","h","ggrt::render_forward_kernel(ggrt::View)","1","7","(256, 1, 1)","(63, 48, 1)","0","10.0","s","gpu__time_duration.sum","ns","83000"\n'
        '"2","1","python","h","ggrt::render_forward_kernel(ggrt::View)","1","7","(256, 1, 1)","(63, 48, 1)","0","10.0","s","gpu__time_duration.sum","ns","85,000"\n')
    la = ns.read_launches(str(launches))
    assert la["render_forward_kernel"] == [83.0, 85.0]
    assert abs(la["at::fill_kernel<float>"][0] - 38.24) < 1e-9


def test_committed_profiles_are_consistent():
    import json

    d = json.loads((ROOT / "profiles" / "r1_ncu_kernels.json").read_text())
    assert d["workload"].startswith("C2")
    shares = d["launch_list_share_of_step"]
    assert abs(sum(shares.values()) - 1.0) < 1e-3 and len(shares) == 8
    line = json.loads((ROOT / "profiles" / "r1_bench_c2.json").read_text())
    for key in ("metric", "value", "unit", "n_gpus", "ms_per_step", "e2e", "roofline", "cpu_baseline", "clocks", "gpu_launches"):
        assert key in line, key
    assert line["roofline"]["bound"] == "hbm" and 0 < line["roofline"]["whole_path"]["frac"] < 1
