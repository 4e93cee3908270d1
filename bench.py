#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the board hot path.

Workload (BASELINE.json configs[1]): 4096 concurrent 19x19 games per GPU, random-policy playouts
(include/elfb200_playout_policy.h) from the empty board to GoState::terminated().  One "step" =
one such batch (about 1.86 M plies).  Metric: moves/sec (plies/sec), whole job over all ranks.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, host cores

`value`   device-timed (CUDA events on the library's stream), inputs resident, max over ranks.
`e2e`     the same metric through the public C-ABI call elfb200_playout() with HOST result
          buffers (launch + D2H of checksum/plies/score/hash every step), wall clock.
`roofline` dominant kernel k_playout: SURVEY 8d algorithmic bytes (step 264 B + legal mask 184 B
          per game-ply) / CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` the compiled reference (oracle/_ref; else the oracle port) on all host cores for a
          bounded sample of the same workload (rank 0, N=1 only).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BOARD = 19
GAMES_PER_GPU = 4096
SEED = 20260922
ALGO_BYTES_PER_PLY = 264 + 184  # step + legal mask, 19x19 (9x9: 2*(32+32)+8 + 32+32+24 = 224)


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture, or None"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref) or the oracle port, all host threads
# --------------------------------------------------------------------------------------------
def effective_cores():
    """host threads this process can really use: affinity mask capped by the cgroup CPU quota"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def cpu_playouts(seconds=None, games_per_thread=None, first_id=10_000_000):
    """Run random-policy playouts on every host core.  Either time-bounded (`seconds`) or a fixed
    number of games per thread.  Returns dict(moves, seconds, cores, kind, games)."""
    from tests import oracles

    cores = effective_cores()
    if oracles.have_ref(BOARD):
        kind = "reference"
        L = oracles.load_ref(BOARD)

        def one(gid):
            chk = ctypes.c_uint64()
            sc = ctypes.c_int32()
            return L.ref_playout(SEED, gid, 2 * BOARD * BOARD, None, None, None, ctypes.byref(chk), ctypes.byref(sc))
    else:
        kind = "port"
        L = oracles.load_oracle()

        def one(gid):
            chk = ctypes.c_uint64()
            sc = ctypes.c_int32()
            return L.go_playout(BOARD, SEED, gid, 2 * BOARD * BOARD, None, None, None, ctypes.byref(chk), ctypes.byref(sc))

    moves = [0] * cores
    games = [0] * cores
    t0 = time.perf_counter()
    deadline = t0 + seconds if seconds else None

    def work(tid):
        gid = first_id + tid * 1_000_000
        n = 0
        while True:
            if deadline is not None and time.perf_counter() >= deadline:
                break
            if games_per_thread is not None and n >= games_per_thread:
                break
            moves[tid] += one(gid + n)  # ctypes releases the GIL during the call
            n += 1
        games[tid] = n

    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return {"moves": sum(moves), "seconds": dt, "cores": cores, "kind": kind, "games": sum(games)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    gpt = 24  # games per thread per step: ~0.15-0.3 s of CPU work per step
    for _ in range(args.warmup):
        cpu_playouts(games_per_thread=4)
    tot_moves, tot_s, info = 0, 0.0, None
    for _ in range(args.steps):
        info = cpu_playouts(games_per_thread=gpt)
        tot_moves += info["moves"]
        tot_s += info["seconds"]
    val = tot_moves / tot_s
    sample = f"{gpt} playouts/thread/step x {info['cores']} threads (of the 4096-game batch), {args.steps} steps"
    line = {
        "impl": "reference", "metric": f"self-play moves/sec (random-policy playouts, {BOARD}x{BOARD})", "value": val,
        "unit": "moves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"configs[1]: {args.games} concurrent {BOARD}x{BOARD} games, random-policy playouts (each host thread plays games back to back)",
                   "games_per_gpu": args.games, "board": BOARD, "seed": SEED},
        "cpu_baseline": {"value": val, "unit": "moves/s", "cores": info["cores"], "kind": info["kind"], "sample": sample},
        "e2e": {"value": val, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_ours(args):
    import torch

    import elf_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    G = args.games
    gb = elf_b200.GoBatch(G, board_size=BOARD, device=local)
    stream = torch.cuda.ExternalStream(gb.stream, device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    from elf_b200.dist_utils import reduce_timing_and_counts, shard_first_game_id

    def first_id(step):  # distinct games per (step, rank)
        return shard_first_game_id(step, world, rank, G)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    PLIES = args.plies_per_slot

    # ---- warm-up -----------------------------------------------------------------------------
    for w in range(args.warmup):
        gb.playout_stream_launch(SEED, first_id(10_000 + w), PLIES)
    gb.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-timed region: K steps, CUDA events on the library's stream, L2 flushed between --
    launches0 = gb.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    plies_total = 0
    barrier()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(s & 0xFF)  # evict L2 (outside the event pair)
            ev[s][0].record(stream)
        gb.playout_stream_launch(SEED, first_id(s), PLIES)
        with torch.cuda.stream(stream):
            ev[s][1].record(stream)
        gb.synchronize()
        plies_total += gb.playout_results()["total_plies"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = gb.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- e2e: the public call with host result buffers, wall clock ------------------------------
    barrier()
    t0 = time.perf_counter()
    e2e_plies = 0
    for s in range(args.steps):
        r = gb.playout_stream(SEED, first_id(s), PLIES)  # launch + D2H(chk, plies, games, hash) + sync
        e2e_plies += r["total_plies"]
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- secondary: one batch of G games played to terminal (includes the ragged tail) -----------
    tt_ms, tt_plies = 0.0, 0
    for s in range(min(args.steps, 10)):
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(stream)
        gb.playout_launch(SEED, first_id(s))
        with torch.cuda.stream(stream):
            b2.record(stream)
        gb.synchronize()
        tt_ms += a.elapsed_time(b2)
        tt_plies += gb.playout_results()["total_plies"]
    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary: the step API (GoState::forward for the whole batch per call, HOST buffers) --------
    step_api = None
    if rank == 0:
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", f"playouts_{BOARD}.json")))
            mv = next(e["moves"] for e in gold["games"] if "moves" in e)
            gb.reset()
            acts = np.empty(G, np.int32)
            gb.forward(np.full(G, mv[0], np.int32))
            gb.reset()
            t_s = time.perf_counter()
            for a in mv:
                acts.fill(a)
                ok = gb.forward(acts)  # H2D actions, k_step, D2H accept flags, sync
            dt_s = time.perf_counter() - t_s
            step_api = {"value": G * len(mv) / dt_s, "unit": "moves/s", "us_per_call": 1e6 * dt_s / len(mv),
                        "all_accepted": bool(ok.all()), "h2d_bytes_per_call": 4 * G, "d2h_bytes_per_call": G,
                        "note": "elfb200_step(): every game replays one reference move list, one call per ply"}
        except Exception as e:
            step_api = f"unmeasured: {e}"

    # ---- spot parity of timed work (rank 0): a few games of step 0 against the oracle ------------
    parity = None
    if rank == 0:
        try:
            from tests import oracles

            r0 = gb.playout_stream(SEED, first_id(0), PLIES)
            ok = True
            for g in (0, 1337, G - 1):
                t, acc, games = oracles.oracle_playout_stream(BOARD, SEED, first_id(0), g, G, PLIES)
                ok &= (t, acc, games) == (int(r0["plies"][g]), int(r0["chk"][g]), int(r0["games"][g]))
            parity = bool(ok)
        except Exception as e:  # oracle missing is not fatal for the bench
            parity = f"unchecked: {e}"

    # ---- reduce over ranks (MAX of times, SUM of counters) ----------------------------------------
    (dev_ms, e2e_s), (plies_total, e2e_plies, launches) = reduce_timing_and_counts(
        dist, f"cuda:{local}", [dev_ms, e2e_s], [plies_total, e2e_plies, launches])

    if rank == 0:
        peak, peak_src = measured_peaks()
        value = plies_total / (dev_ms / 1e3)
        per_rank_plies = plies_total / world
        algo = ALGO_BYTES_PER_PLY if BOARD == 19 else 224
        achieved = algo * per_rank_plies / args.steps / (dev_ms / args.steps / 1e3) / 1e9
        line = {
            "metric": f"self-play moves/sec (random-policy playouts, {BOARD}x{BOARD})", "value": value, "unit": "moves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": (f"configs[1]: 4096 concurrent 19x19 games per GPU" if (BOARD, G) == (19, 4096) else f"configs[4]-style: {G} concurrent {BOARD}x{BOARD} games per GPU") + f", random-policy playouts, steady state: every game slot plays {PLIES} plies per step and restarts finished games",
                       "games_per_gpu": G, "board": BOARD, "seed": SEED, "plies_per_step": plies_total / args.steps,
                       "plies_per_slot": PLIES,
                       "l2": "flushed (256 MiB write) between timed steps", "parallelism": f"games sharded x{world}, no collective"},
            "e2e": {"value": e2e_plies / e2e_s, "unit": "moves/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 24 * G,
                    "note": "elfb200_playout_stream() from the host, results read back every step; this workload's inputs are 3 scalars "
                            "(seed, first id, plies per slot) passed as kernel params, so there is nothing to copy in.  The host-driven "
                            "flavour of the same path -- one elfb200_step() per ply with HOST action and accept buffers -- is in host_driven",
                    "host_driven": step_api if isinstance(step_api, dict) else {"unmeasured": str(step_api)}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic("k_playout<19>") if BOARD == 19 else None, "peak_source": peak_src,
                         "kernel": f"k_playout<{BOARD}>", "algorithmic_bytes_per_ply": algo,
                         "note": "position and group masks live in registers, the superko record in L2: DRAM is idle and the kernel is bound by the integer ALU pipe (profiles/r1_playout_F.md)"},
            "clocks": clocks, "wall_s_timed_region": t_wall, "parity_spot_check": parity, "step_api": step_api,
            "batch_to_terminal": {"value": tt_plies / (tt_ms / 1e3) * 1.0, "unit": "moves/s (this rank)",
                                  "ms_per_batch": tt_ms / max(1, min(args.steps, 10)),
                                  "note": "one batch of 4096 games from the empty board to terminated(): includes the ragged tail"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_playouts(seconds=args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": cb["moves"] / cb["seconds"], "unit": "moves/s", "cores": cb["cores"], "kind": cb["kind"],
                "sample": f"{cb['games']} playouts of the same workload in {cb['seconds']:.1f} s on {cb['cores']} threads"}
        emit(line)
    gb.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------
# secondary workload: BASELINE configs[2] -- MCTS self-play (the network is PyTorch/cuDNN plumbing;
# our kernels are select / leaf features / expand / backup).  `--workload mcts`.
# --------------------------------------------------------------------------------------------
def cpu_mcts_rollouts(seconds, rollouts, per_batch):
    """reference TreeSearchT (oracle/_ref) with the deterministic fake net on every host thread:
    rollouts/s without any network cost."""
    from tests import oracles

    cores = effective_cores()
    if not oracles.have_ref(BOARD):
        return None
    done = [0] * cores
    t0 = time.perf_counter()
    deadline = t0 + seconds

    def work(tid):
        st = oracles.Ref(BOARD)
        m = oracles.RefMcts(BOARD, num_rollouts=rollouts, num_rollouts_per_batch=per_batch, virtual_loss=1,
                            persistent_tree=1, c_puct=1.5, seed=tid)
        while time.perf_counter() < deadline:
            r = m.act(st)
            st.forward(r["best_action"])
            done[tid] += 1
            if st.terminated():
                break

    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return {"moves": sum(done), "seconds": dt, "cores": cores}


def cpu_mcts_with_net(seconds, rollouts, per_batch, actor, device):
    """the reference search (oracle/_ref, one search thread per game, one game
    per host thread) driving the SAME GPU network through a callback -- what the reference's own
    batching can extract from the net when it only has the host cores to run the search on."""
    import torch

    from tests import oracles

    cores = effective_cores()
    if not oracles.have_ref(BOARD):
        return None
    done = [0] * cores
    evals = [0] * cores
    t0 = time.perf_counter()
    deadline = t0 + seconds

    def work(tid):
        def cb(feats, hashes):
            with torch.no_grad():
                out = actor({"s": torch.from_numpy(np.ascontiguousarray(feats)).to(device)})
            evals[tid] += len(hashes)
            return out["pi"].float().cpu().numpy(), out["V"].float().reshape(-1).cpu().numpy()

        st = oracles.Ref(BOARD)
        m = oracles.RefMcts(BOARD, num_rollouts=rollouts, num_rollouts_per_batch=per_batch, virtual_loss=1,
                            persistent_tree=1, c_puct=1.5, seed=tid, callback=cb)
        while time.perf_counter() < deadline and not st.terminated():
            r = m.act(st)
            st.forward(r["best_action"])
            done[tid] += 1

    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return {"moves": sum(done), "seconds": dt, "cores": cores, "evals": sum(evals)}


def run_mcts(args):
    import torch

    import elf_b200
    from elf_b200.model import Actor, PolicyValueNet, broadcast_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G, R, B = args.games, args.rollouts, args.per_batch
    dev = torch.device("cuda", local)
    torch.manual_seed(1234)
    if args.fake_net:
        P1 = BOARD * BOARD + 1
        table = torch.rand(4096, P1, device=dev).softmax(1)
        vals = torch.rand(4096, device=dev) * 2 - 1

        def actor(batch):
            n = batch["s"].shape[0]
            idx = torch.arange(n, device=dev) % 4096
            return {"pi": table[idx], "V": vals[idx]}
        net_desc = "fake (table lookup, engine-only timing)"
    else:
        model = PolicyValueNet(BOARD, num_block=args.blocks, dim=args.dim).to(dev)
        t_b0 = time.perf_counter()
        broadcast_weights(model)  # frozen weights from rank 0: the only collective of the path
        torch.cuda.synchronize()
        t_bcast = time.perf_counter() - t_b0
        actor = Actor(model, batchsize=args.nn_batch)
        net_desc = f"random-init resnet {args.blocks}x{args.dim}, fp16 weights, channels_last, NN batch {args.nn_batch}"
    sp = elf_b200.selfplay.SelfPlay(actor, num_games=G, board_size=BOARD, device=local, policy_distri_cutoff=0,
                                    resign_thres=0.0, never_resign_ratio=1.0, num_rollouts=R,
                                    num_rollouts_per_batch=B, virtual_loss=1, persistent_tree=1, c_puct=1.5,
                                    rotation_flip=1, seed=rank)
    ext = torch.cuda.ExternalStream(sp.gb.stream, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(ext):
        for _ in range(args.warmup):
            sp.step()
        sp.mcts.timings(reset=True)
        st0 = sp.mcts.stats().astype(np.int64)
        ev0 = sp.mcts.eval_count()
        l0 = sp.gb.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        t0 = time.perf_counter()
        moves = 0
        for _ in range(args.steps):
            moves += sp.step()
        e1.record(ext)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        ms, waves = sp.mcts.timings()
        st = sp.mcts.stats().astype(np.int64) - st0
        evals = sp.mcts.eval_count() - ev0
        launches = sp.gb.launch_count() - l0
    if dist is not None:
        t = torch.tensor([dev_ms, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = t.tolist()
        c = torch.tensor([moves, launches], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        moves, launches = c.tolist()
    if rank == 0:
        peak, peak_src = measured_peaks()
        E = BOARD * BOARD + 1
        sel_bytes = int(st[0]) * (32 + 4) + int(st[3]) * 16  # header + E_n*16 + vl write (full scan)
        sel_bytes_read = int(st[0]) * (32 + 4) + int(st[1]) * 16  # what the prefix scan actually reads
        feat_bytes = evals * 26792
        sel_gbs = sel_bytes / (ms[0] / 1e3) / 1e9 if ms[0] > 0 else 0.0
        feat_gbs = feat_bytes / (ms[1] / 1e3) / 1e9 if ms[1] > 0 else 0.0
        line = {
            "metric": "self-play moves/sec (MCTS, 19x19)", "value": moves / (dev_ms / 1e3), "unit": "moves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 search / fp16 net",
            "data": "synthetic",
            "config": {"workload": f"configs[2]-shaped: {G} games/GPU x {R} rollouts/move, {B} rollouts/wave, puct 1.5, vloss 1, persistent tree",
                       "net": net_desc, "games_per_gpu": G, "rollouts": R, "l2": "node pool >> L2 (7.4 KB/node)",
                       "parallelism": f"games sharded x{world}, NCCL weight broadcast only"},
            "e2e": {"value": moves / wall, "unit": "moves/s", "h2d_bytes_per_step": 4 * G + G,
                    "d2h_bytes_per_step": int(G * (E * 4 + 16 + 48 * 2)),
                    "note": "through SelfPlay.step(): actions H2D, root tables + info D2H each move; leaf features never leave the GPU"},
            "gpu_launches": int(launches),
            "kernels_ms_per_wave": {"select": ms[0] / max(waves, 1), "leaf_features": ms[1] / max(waves, 1),
                                    "expand": ms[2] / max(waves, 1), "backup": ms[3] / max(waves, 1)},
            "waves": int(waves), "nn_evals": int(evals), "rollouts_per_s": (int(waves) * B * G * world) / (dev_ms / 1e3),
            "roofline": {"bound": "hbm", "achieved": sel_gbs, "peak": peak, "unit": "GB/s", "frac": sel_gbs / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "k_select<19>",
                         "algorithmic_bytes": sel_bytes, "bytes_read_by_prefix_scan": sel_bytes_read,
                         "nodes_visited": int(st[0]), "edges_scanned": int(st[1]), "edges_stored": int(st[3]),
                         "note": "achieved uses the SURVEY 8d full-scan byte formula; the kernel reads only the selected prefix of each node's edges (same arg-max), so achieved can exceed what DRAM delivers",
                         "leaf_features_GBps": feat_gbs},
            "clocks": clocks,
        }
        if not args.fake_net:
            line["weight_broadcast_s"] = t_bcast
        if world == 1 and not args.no_cpu_baseline and not args.fake_net:
            cn = cpu_mcts_with_net(min(args.cpu_seconds, 20.0), R, B, actor, dev)
            if cn:
                line["cpu_baseline_same_net"] = {
                    "value": cn["moves"] / cn["seconds"], "unit": "moves/s", "cores": cn["cores"], "kind": "reference",
                    "nn_positions_per_s": cn["evals"] / cn["seconds"],
                    "sample": f"reference TreeSearchT on {cn['cores']} host threads (one game each, {B} leaves per NN call) driving the same GPU network: {cn['moves']} moves in {cn['seconds']:.1f} s"}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_mcts_rollouts(min(args.cpu_seconds, 15.0), R, B)
            if cb:
                line["cpu_baseline"] = {"value": cb["moves"] / cb["seconds"], "unit": "moves/s", "cores": cb["cores"],
                                        "kind": "reference",
                                        "sample": f"reference TreeSearchT, {R} rollouts/move, 1 search thread per game, fake net (no NN cost), {cb['moves']} moves in {cb['seconds']:.1f} s on {cb['cores']} threads"}
        emit(line)
    sp.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else any library prints (e.g. the
    'NCCL version' banner) was re-routed to stderr in main()"""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT, BOARD
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # fd 1 -> stderr for native libraries and stray prints
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--plies-per-slot", type=int, default=512)
    ap.add_argument("--workload", default="playout", choices=["playout", "mcts"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--board", type=int, default=19, choices=[9, 19])
    ap.add_argument("--rollouts", type=int, default=64)
    ap.add_argument("--per-batch", type=int, default=8)
    ap.add_argument("--blocks", type=int, default=20)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--nn-batch", type=int, default=256)
    ap.add_argument("--fake-net", action="store_true")
    args = ap.parse_args()
    BOARD = args.board
    if args.workload == "mcts" and args.impl == "ours":
        return run_mcts(args)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
