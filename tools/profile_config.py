import os, sys, tempfile, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import config_runs
REF = config_runs.find_reference()
key = sys.argv[1] if len(sys.argv) > 1 else "C4"
seeds = sys.argv[2] if len(sys.argv) > 2 else None      # e.g. 0-4: profile the in-process seed loop
cfg = config_runs.CONFIGS[key]
with tempfile.TemporaryDirectory() as work:
    config_runs.prepare_workdir(work, REF, key)
    script = os.path.join(REF, cfg["script"])
    args = ["--seed", "0", "--config", cfg["cfg"], "--gpu", "0"] + cfg["extra"]
    if seeds:
        args = [a for i, a in enumerate(args) if a != "--seed" and (i == 0 or args[i - 1] != "--seed")]
    cmd = [sys.executable, "-m", "cProfile", "-o", os.path.join(work, "prof.out"), config_runs.LAUNCHER, "--reference", REF] + \
        (["--seeds", seeds] if seeds else []) + [script] + args
    t0 = time.time()
    r = subprocess.run(cmd, cwd=work, env=config_runs._env({"SINDY_B200_INIT_RNG": "cpu"}), capture_output=True, text=True)
    print("wall", time.time() - t0, "rc", r.returncode, r.stderr[-500:])
    import pstats
    p = pstats.Stats(os.path.join(work, "prof.out"))
    p.sort_stats("cumulative").print_stats(45)
