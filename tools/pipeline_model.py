"""Discrete-event model of the multi-GPU schedules of slab.py (no GPU needed): N ranks, G groups of right-hand sides.

One Krylov iteration of a group = an application of the preconditioner (forward slab sweeps on ranks 0..N-1 one after the
other, backward sweeps on ranks N-1..0) followed by its vector phase (SpMV + Gram-Schmidt: `syncs` sub-steps, each a short
piece of work on EVERY rank followed by a global synchronisation, the all-reduce).  A rank runs one thing at a time (a slab
sweep holds every cluster slot of the GPU; vector kernels queue behind it) in the order the work becomes ready.

    lockstep   all groups do their forward sweeps, then all their backward sweeps, then all their vector phases
    async      every group runs its own loop (slab.GroupPipeline); a rank serves whatever is ready first

Parameters are the measured single-GPU numbers of bench.py at 4096^2 with 8 right-hand sides per group:
    sweep_ms   one full sweep of a group (31.7 ms), a slab sweep is sweep_ms / N + handover_ms
    applies    preconditioner applications per inner iteration (22 per restart cycle of 20 = 1.1)
    vector_ms  vector work of one group per iteration on ONE GPU (19.2 ms); 1/N of it per rank
    syncs      global synchronisations per iteration of a group (k + 2 all-reduces of MGS, k = 10.5 on average, + norms, halo)
    sync_ms    latency of one synchronisation (all-reduce + the host round trip that follows some of them)

usage: python tools/pipeline_model.py            -> table of modelled against measured ms per step
"""
import heapq
import sys


def simulate(N, G, mode, iters=12, sweep_ms=31.7, applies=1.1, vector_ms=19.2, syncs=14, sync_ms=0.06, handover_ms=0.08):
    """returns ms per step (= one inner iteration of every group) in the steady state"""
    slab = sweep_ms / N + handover_ms
    vpiece = vector_ms / N / syncs
    free = [0.0] * N                                   # time at which rank r is free again

    def run(r, ready, dur):
        start = max(free[r], ready)
        free[r] = start + dur
        return free[r]

    def apply_M(t):                                    # one group, starting at time t on rank 0; returns completion time
        for r in range(N):
            t = run(r, t, slab)
        for r in range(N - 1, -1, -1):
            t = run(r, t, slab)
        return t

    def vector_phase(t):
        for _ in range(syncs):
            t = max(run(r, t, vpiece) for r in range(N)) + sync_ms
        return t

    if mode == "lockstep":
        t_done = [0.0] * G
        marks = []
        for it in range(iters):
            # forward chains of all groups (group g enters rank 0 after group g-1), then the backward chains, then vectors
            n_apply = 2 if (it % 10 == 0) else 1       # 22 applications per 20 iterations
            for _ in range(n_apply):
                fw = []
                for g in range(G):
                    t = t_done[g]
                    for r in range(N):
                        t = run(r, t, slab)
                    fw.append(t)
                for g in range(G):
                    t = fw[g]
                    for r in range(N - 1, -1, -1):
                        t = run(r, t, slab)
                    t_done[g] = t
            t0 = max(t_done)
            t = t0
            for g in range(G):                         # lock step: the vector phases of the groups one after the other
                t = vector_phase(t)
            t_done = [t] * G
            marks.append(t)
        return (marks[-1] - marks[1]) / (len(marks) - 2)
    # async: event-driven; every group is a state machine, a rank serves requests in the order they become ready
    # states per group: list of (kind, rank) operations of one iteration
    def ops_of_iteration(it):
        ops = []
        n_apply = 2 if (it % 10 == 0) else 1
        for _ in range(n_apply):
            ops += [("S", r) for r in range(N)] + [("S", r) for r in range(N - 1, -1, -1)]
        ops += [("V", None)] * syncs
        return ops
    # process events in time order so that FIFO-by-readiness holds approximately
    pq = []                                            # (ready time, seq, group)
    prog = {g: (0, 0) for g in range(G)}               # (iteration, index of the next operation)
    seq = 0
    for g in range(G):
        heapq.heappush(pq, (0.0, seq, g)); seq += 1
    finish = {g: [] for g in range(G)}
    cache = {}
    while pq:
        t, _, g = heapq.heappop(pq)
        it, i = prog[g]
        if it >= iters:
            continue
        if it not in cache:
            cache[it] = ops_of_iteration(it)
        kind, r = cache[it][i]
        if kind == "S":
            t2 = run(r, t, slab)
        else:
            t2 = max(run(q, t, vpiece) for q in range(N)) + sync_ms
        i += 1
        if i == len(cache[it]):
            finish[g].append(t2)
            it, i = it + 1, 0
        prog[g] = (it, i)
        heapq.heappush(pq, (t2, seq, g)); seq += 1
    # steady state: average time per iteration of a group over the last iterations = time per step
    per = [(f[-1] - f[1]) / (len(f) - 2) for f in finish.values()]
    return sum(per) / len(per)


MEASURED = {  # ms per step, profiles/r02c_bench_n*.json (10 steps); N = 1: 88.7
    (2, "lockstep"): 130.3, (2, "async"): 121.6, (4, "async"): 112.9, (8, "lockstep"): 171.0, (8, "async"): 136.2,
}


def main():
    print("| GPUs | schedule | modelled ms/step | measured ms/step | modelled efficiency |")
    print("|---|---|---|---|---|")
    one = simulate(1, 1, "lockstep")
    for N in (2, 4, 8):
        for mode in ("lockstep", "async"):
            ms = simulate(N, N, mode)
            meas = MEASURED.get((N, mode))
            print(f"| {N} | {mode} | {ms:.1f} | {meas if meas else '-'} | {100 * one / ms:.0f} % |")
    print(f"\n(one GPU, one group: {one:.1f} ms per step modelled, 88.7 measured)")
    if "--sweep" in sys.argv:
        print("\nasync at N = 8: what the levers would buy (ms per step)")
        for name, kw in (("as measured", {}), ("2 groups per GPU", {"G": 16}), ("3 synchronisations per iteration (block Gram-Schmidt at MGS bandwidth)", {"syncs": 5}),
                         ("no hand-over latency", {"handover_ms": 0.0}), ("vector phase twice as fast", {"vector_ms": 9.6})):
            G = kw.pop("G", 8)
            print(f"  {name}: {simulate(8, G, 'async', **kw) * 8 / G:.1f}")


if __name__ == "__main__":
    main()
