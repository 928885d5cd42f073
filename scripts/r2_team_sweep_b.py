"""Round-2 GPU call B: k_team after the instruction diet."""
import json, os, subprocess, sys, time
OUT = os.path.join("gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "r2b")
os.makedirs(OUT, exist_ok=True)
T = "ubench/team_probe"
log = open(os.path.join(OUT, "team.jsonl"), "a")


def run(args, timeout=60):
    t0 = time.time()
    try:
        r = subprocess.run([T] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
        rec = None
        for line in r.stdout.strip().splitlines():
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


# args: C ns mode log2n groups [budget] [cap] [reps] [max_clusters] [dist]
for C in (2, 4):
    for ns in (0, 1):
        for M in (0, 1):
            for dist in (0, 1, 2):
                run([C, ns, M, 22, 100001, 199552, 0, 2, 0, dist], timeout=40)
    run([C, 0, 1, 22, 100000, 199552, 8, 2, 3, 0], timeout=60)
for C in (2, 4):
    for M in (0, 1):
        run([C, 0, M, 30, 100000, 199552, 0, 5, 0, 0], timeout=90)
for C, G, budget, cap in ((2, 60000, 199552, 0), (2, 76000, 199552, 0), (2, 100000, 199552, 72), (2, 100000, 199552, 64), (2, 100000, 215000, 0),
                          (2, 100000, 232000, 0), (4, 150000, 199552, 0), (4, 100000, 199552, 40), (2, 30000, 199552, 0), (2, 1000000, 199552, 0)):
    run([C, 0, 0, 30, G, budget, cap, 5, 0, 0], timeout=90)
for G in (60000, 100000):
    run([2, 0, 1, 30, G, 199552, 0, 5, 0, 0], timeout=90)
    run([4, 0, 1, 30, G, 199552, 0, 5, 0, 0], timeout=90)
    run([4, 0, 1, 30, G, 226000, 0, 5, 0, 0], timeout=90)
