"""CPU: bench.py's contract pieces that need no GPU — argument defaults per --gpus N, the config block, and the reference arm's JSON line
(the oracle port on the host cores, one frame per step) at --gpus 1 and, as a non-zero rank of a torchrun launch, the silent exit."""
import argparse
import importlib.util
import json
import os
import subprocess
import sys

from conftest import ROOT


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def args_for(**kw):
    base = dict(gpus=1, steps=10, warmup=3, impl="ours", config=None, no_color=False, no_mc=False, cpu_frames=24, no_cpu_baseline=False, ray_steps=0,
                frames_per_step=0, pool_blocks=0, all_frames=False, no_ref_cuda=False, no_c4=False)
    base.update(kw)
    return argparse.Namespace(**base)


def test_defaults_follow_the_gpu_count(synth):
    b = load_bench()
    a1 = args_for(gpus=1); b.apply_defaults(a1)
    assert a1.config == "C2" and b.METRIC == "depth_frames_per_sec_640x480_5mm" and b.FRAMES_PER_STEP == 50 and not a1.ray_steps
    b = load_bench()
    a8 = args_for(gpus=8); b.apply_defaults(a8)
    assert a8.config == "C4" and a8.ray_steps == 1100 and a8.frames_per_step == 4 and a8.pool_blocks == 16 << 20
    assert b.METRIC == "depth_frames_per_sec_640x480_2mm_room_scale" and b.FRAMES_PER_STEP == 4
    cfg, sc = synth.CONFIGS["C4"], synth.make_scene("C4", color=True)
    assert b.frame_cap(a8, sc) == 16                                  # what one GPU holds; --all-frames lifts it
    a8.all_frames = True
    assert b.frame_cap(a8, sc) == sc.n_frames
    a8.all_frames = False
    c = b.config_dict(a8, cfg, sc, 8)
    assert "config 4" in c["workload"] and "16 frames" in c["workload"] and c["max_ray_steps"] == 1100 and "replayed" in c["sequence_replay"]
    assert "L2" in c["l2"]


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")                           # what torchrun exports: the arm must set its thread count itself
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and len(lines) == 1, out.stdout[-1000:] + out.stderr[-2000:]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "depth_frames_per_sec_640x480_5mm" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ncpu = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == ncpu, "the arm must use every host core even when OMP_NUM_THREADS=1 is inherited"


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == "", out.stdout[-500:] + out.stderr[-500:]
