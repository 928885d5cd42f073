"""Round-2 GPU call A: transport microbenchmark, correctness and timing sweep of k_team (ubench/team_probe).
Every case runs in its own process under a timeout (a protocol bug would show up as a hang)."""
import json, os, subprocess, sys, time

OUT = os.path.join("gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "r2a")
os.makedirs(OUT, exist_ok=True)
T = "ubench/team_probe"
log = open(os.path.join(OUT, "team.jsonl"), "a")
broken = set()


def run(args, timeout=60, key=None):
    if key is not None and key in broken:
        return None
    t0 = time.time()
    try:
        r = subprocess.run([T] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
        out = r.stdout.strip().splitlines()
        rec = None
        for line in out:
            try:
                rec = json.loads(line)
            except Exception:
                pass
        if rec is None:
            rec = {"failed": args, "rc": r.returncode, "stdout": r.stdout[-300:], "stderr": r.stderr[-300:]}
    except subprocess.TimeoutExpired:
        rec = {"failed": args, "timeout": timeout}
    rec["wall_s"] = round(time.time() - t0, 2)
    log.write(json.dumps(rec) + "\n"); log.flush()
    print(json.dumps(rec), flush=True)
    ok = "failed" not in rec and "error" not in rec and rec.get("max_rel_err", 0) < 1e-9 and rec.get("count_mismatch", 0) == 0
    if not ok and key is not None:
        broken.add(key)
    return rec


# 1. transport only
for C, X, K in ((2, 0, 64), (2, 1, 64), (4, 0, 32), (4, 1, 32), (8, 0, 16)):
    run(["xport", C, X, K, 4000], timeout=40)
# 2. correctness at N = 2^22 (+3 ragged), 10^5 groups: uniform, skewed, signed values with tiny ones
for C in (2, 4):
    for X in (0, 1):
        for M in (0, 1):
            for dist in (0, 1, 2):
                run([C, X, M, 22, 100000, 199552, 0, 2, 0, dist], timeout=40, key=(C, X))
# tiny buffers force the overflow path; few clusters force many phases per warp
for C in (2, 4):
    run([C, 0, 1, 22, 100000, 199552, 8, 2, 3, 0], timeout=60, key=(C, 0))
# 3. timing
for log2n in (28, 30):
    for C in (2, 4):
        for X in (0, 1):
            for M in (0, 1):
                run([C, X, M, log2n, 100000, 199552, 0, 5, 0, 0], timeout=90, key=(C, X))
# 4. shared-memory budget / group-count variants at N = 2^30 (sum)
for C, G, budget in ((2, 100000, 232000), (2, 76000, 199552), (2, 60000, 199552), (4, 150000, 199552), (4, 190000, 199552), (4, 100000, 232000), (8, 100000, 199552)):
    run([C, 0, 0, 30, G, budget, 0, 5, 0, 0], timeout=90, key=(C, 0))
