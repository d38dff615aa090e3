#!/usr/bin/env python
"""Runs the reference's own conformance programs -- rdp_conformance.cpp / vi_conformance.cpp compiled unmodified from
/root/reference into oracle/_ref/{rdp,vi}-conformance-b200 with oracle/ref_harness/replayer_driver_b200.cpp in the GPU
seat (see oracle/Makefile) -- over the suites and iteration ranges the reference registers with ctest
(tests/golden/reference_ctest_list.txt: 231 rdp + 25 vi tests, --range 0 1000, 100 for the many-primitive ones).

    python tools/run_reference_conformance.py [--library PATH] [--jobs N] [--scale F] [--filter SUBSTR] [--log FILE]

--library  the rdp library the harness loads (default: the product, parallel-rdp_b200/librdp_b200.so; the CPU suite
           passes the test-only host simulator)
--scale    fraction of each registered iteration count to run (1.0 = the reference's full depth)
Every suite is one process (the oracle is a process-wide singleton); --jobs of them run at a time.
"""
import argparse
import concurrent.futures
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIST = os.path.join(ROOT, "tests", "golden", "reference_ctest_list.txt")
BIN = {"rdp": os.path.join(ROOT, "oracle", "_ref", "rdp-conformance-b200"), "vi": os.path.join(ROOT, "oracle", "_ref", "vi-conformance-b200")}


def registered_tests():
    out = []
    for line in open(LIST):
        if line.strip() and not line.startswith("#"):
            prog, name, n = line.split()
            out.append((prog, name, int(n)))
    return out


def regenerate_list(reference="/root/reference"):
    txt = open(os.path.join(reference, "CMakeLists.txt")).read()
    rows = []
    for m in re.finditer(r"^\s*add_(rdp_test_reduced|rdp_test|vi_test)\(([^)]+)\)", txt, re.M):
        kind, name = m.group(1), m.group(2).strip()
        rows.append(("vi" if kind == "vi_test" else "rdp", name, 100 if kind == "rdp_test_reduced" else 1000))
    head = [l for l in open(LIST) if l.startswith("#")]
    with open(LIST, "w") as f:
        f.writelines(head)
        f.writelines(f"{a} {b} {c}\n" for a, b, c in rows)
    return len(rows)


def available():
    return all(os.path.exists(p) for p in BIN.values())


def run_one(prog, name, iterations, library, timeout):
    env = dict(os.environ)
    if library:
        env["RDP_B200_LIBRARY"] = library
    t0 = time.time()
    try:
        r = subprocess.run([BIN[prog], "--suite", name, "--range", "0", str(iterations)], capture_output=True, text=True, env=env, timeout=timeout)
        text = r.stdout + r.stderr
        ok = r.returncode == 0 and "====== PASSED ======" in text
        # the suite has to have run: a name the binary does not know is skipped silently
        ran = f"Running suite: {name}" in text or ok
        tail = "\n".join(l for l in text.splitlines() if "Skipping" not in l)[-600:]
    except subprocess.TimeoutExpired:
        ok, ran, tail = False, True, "timeout"
    return prog, name, iterations, ok and ran, time.time() - t0, tail


def run_single_process(prog, iterations, library, timeout):
    """All suites of one program in ONE process (how the reference's binaries run without --suite): one CUDA context, one
    pair of processors.  Returns [(suite, ok)] in run order; the program stops at the first failing suite."""
    env = dict(os.environ)
    if library:
        env["RDP_B200_LIBRARY"] = library
    try:
        r = subprocess.run([BIN[prog], "--range", "0", str(iterations)], capture_output=True, text=True, env=env, timeout=timeout)
        text = r.stdout + r.stderr
    except subprocess.TimeoutExpired as e:
        text = ((e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")) + "\n[ERROR]: timeout\n"
    results, current = [], None
    for line in text.splitlines():
        m = re.search(r"Running suite: (\S+)", line)
        if m:
            current = m.group(1)
        elif "====== PASSED ======" in line and current:
            results.append((current, True, ""))
            current = None
    if current:
        tail = "\n".join(l for l in text.splitlines() if "Skipping" not in l and "completely zero" not in l)[-800:]
        results.append((current, False, tail))
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--library", default=None)
    # Processes sharing one GPU are time-sliced: more than a few of them make every synchronisation wait for its turn
    # (measured on B200: 14 jobs -> 0.13 s per iteration and ~50 s of start-up per process; 3 jobs -> milliseconds).
    ap.add_argument("--jobs", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--filter", default="")
    ap.add_argument("--log", default=None)
    ap.add_argument("--timeout", type=float, default=3600.0)
    ap.add_argument("--skip-passed-in", default=None, help="log of an earlier run: tests it lists as PASS are not run again")
    ap.add_argument("--regenerate-list", action="store_true")
    ap.add_argument("--single-process", type=int, default=0, metavar="N",
                    help="run every suite of each program in one process with --range 0 N (one CUDA context instead of 255: "
                         "processes sharing a GPU are time-sliced and every synchronous call then waits for its turn)")
    args = ap.parse_args()
    if args.regenerate_list:
        print(regenerate_list(), "tests listed")
        return 0
    if not available():
        print("oracle/_ref conformance programs are not built (make -C oracle, needs /root/reference)")
        return 2
    if args.single_process:
        t0 = time.time()
        registered = {(p, n) for p, n, _ in registered_tests()}
        lines, failed, passed = [], [], 0
        for prog in ("rdp", "vi"):
            for name, ok, tail in run_single_process(prog, args.single_process, args.library, args.timeout):
                tag = "" if (prog, name) in registered else " (not registered with ctest)"
                lines.append(f"{'PASS' if ok else 'FAIL'} {prog}-test-{name} --range 0 {args.single_process}{tag}")
                print(lines[-1], flush=True)
                if ok and (prog, name) in registered:
                    passed += 1
                if not ok:
                    failed.append(name)
                    print(tail)
                    lines.append(tail)
        summary = (f"{passed} / {len(registered)} reference conformance tests passed, {args.single_process} iterations each, one process per "
                   f"program ({time.time() - t0:.0f} s, library {args.library or 'product'})")
        print(summary)
        if args.log:
            with open(args.log, "w") as f:
                f.write(summary + "\n" + "\n".join(lines) + "\n")
        return 1 if failed or passed != len(registered) else 0
    tests = [(p, n, max(1, int(round(c * args.scale)))) for p, n, c in registered_tests() if args.filter in n]
    if args.skip_passed_in:
        done = set(re.findall(r"^PASS (?:rdp|vi)-test-(\S+) --range", open(args.skip_passed_in).read(), re.M))
        tests = [t for t in tests if t[1] not in done]
    t0 = time.time()
    failed, lines = [], []
    with concurrent.futures.ThreadPoolExecutor(args.jobs) as pool:
        for prog, name, n, ok, dt, tail in pool.map(lambda t: run_one(*t, args.library, args.timeout), tests):
            line = f"{'PASS' if ok else 'FAIL'} {prog}-test-{name} --range 0 {n} ({dt:.1f} s)"
            lines.append(line)
            print(line, flush=True)
            if not ok:
                failed.append(name)
                print(tail)
                lines.append(tail)
    summary = (f"{len(tests) - len(failed)} / {len(tests)} reference conformance tests passed at scale {args.scale} "
               f"({sum(t[2] for t in tests)} iterations, {time.time() - t0:.0f} s, library {args.library or 'product'})")
    print(summary)
    if args.log:
        with open(args.log, "w") as f:
            f.write(summary + "\n" + "\n".join(lines) + "\n")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
