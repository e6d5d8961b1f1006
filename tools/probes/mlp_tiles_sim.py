"""Design aid for the fused-MLP tile list (gemm_tcgen05.cu: mlp_fused_kernel): builds the list the way
build_mlp_tiles() does and replays it with per-tile durations and the two dependency kinds (c_proj waits for the
pair-row's c_fc tiles; c_fc's store waits for the ring slot's previous pair-row) to get makespan / ideal and the
largest ring occupancy.  python tools/probes/mlp_tiles_sim.py [P nfc nproj U lag extra ring]"""
import sys
import heapq


def build(P, nfc, nproj, U, lag, extra):
    tiles = []
    f = 0
    j = 0
    F, J = P * nfc, P * nproj
    fc_last_round = [None] * P
    base = max(1, round(U * nproj * 3.6 / (nfc + nproj * 3.6) / 3.6 * 1.0))  # units on c_proj per round at steady state
    base = max(1, U * nproj // (nfc + nproj))
    r = 0
    start = 0
    while f < F or j < J:
        row = [None] * U
        # how many c_proj tiles are available this round
        avail = 0
        jj = j
        while jj < J and fc_last_round[jj // nproj] is not None and fc_last_round[jj // nproj] <= r - lag:
            avail += 1
            jj += 1
        backlog = avail
        quota = min(avail, base + (extra if backlog > base else 0))
        if f >= F:
            quota = min(avail, U)
        chosen = set((start + k) % U for k in range(quota))
        start = (start + quota) % U
        for u in range(U):
            if u in chosen and j < J:
                row[u] = (1, j // nproj, j % nproj)
                j += 1
            elif f < F:
                pr, n = f // nfc, f % nfc
                row[u] = (0, pr, n)
                f += 1
                if n == nfc - 1:
                    fc_last_round[pr] = r
            elif f >= F and j < J and False:
                pass
        tiles.append(row)
        r += 1
        if r > 10000:
            raise RuntimeError("no progress")
    return tiles


def simulate(tiles, P, nfc, nproj, U, ring, t_fc=1.0, t_proj=3.6):
    # each unit runs its column in order; c_proj(pr) starts after all c_fc(pr,*) ended; c_fc(pr) END (store) waits for
    # all c_proj(pr - ring) started+mainloop (approximate: ended)
    R = len(tiles)
    pos = [0] * U
    t = [0.0] * U
    fc_done_cnt = [0] * P
    fc_done_time = [0.0] * P
    pj_done_cnt = [0] * P
    pj_done_time = [0.0] * P
    done = 0
    total = sum(1 for row in tiles for x in row if x)
    # event-driven: repeatedly pick the unit with the smallest time whose next tile's dependencies are resolved
    stall = 0.0
    max_occ = 0
    while done < total:
        progressed = False
        order = sorted(range(U), key=lambda u: t[u])
        for u in order:
            while pos[u] < R and tiles[pos[u]][u] is None:
                pos[u] += 1
            if pos[u] >= R:
                continue
            typ, pr, n = tiles[pos[u]][u]
            if typ == 1:
                if fc_done_cnt[pr] < nfc:
                    continue
                s = max(t[u], fc_done_time[pr])
                stall += s - t[u]
                t[u] = s + t_proj
                pj_done_cnt[pr] += 1
                pj_done_time[pr] = max(pj_done_time[pr], t[u])
            else:
                if pr >= ring and pj_done_cnt[pr - ring] < nproj:
                    continue
                s = t[u]
                e = s + t_fc
                if pr >= ring:
                    e2 = max(e, pj_done_time[pr - ring])
                    stall += e2 - e
                    e = e2
                t[u] = e
                fc_done_cnt[pr] += 1
                fc_done_time[pr] = max(fc_done_time[pr], e)
            pos[u] += 1
            done += 1
            progressed = True
            break
        if not progressed:
            raise RuntimeError("deadlock")
    ideal = (P * nfc * t_fc + P * nproj * t_proj) / U
    return max(t), ideal, stall


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    P, nfc, nproj, U, lag, extra, ring = (a + [197, 12, 3, 74, 3, 4, 32][len(a):])
    tiles = build(P, nfc, nproj, U, lag, extra)
    per_unit = [sum(1 for row in tiles if row[u] and row[u][0] == 1) for u in range(U)]
    print("rounds", len(tiles), "proj per unit min/max", min(per_unit), max(per_unit))
    mk, ideal, stall = simulate(tiles, P, nfc, nproj, U, ring)
    print("makespan %.1f ideal %.1f eff %.3f stall %.1f" % (mk, ideal, ideal / mk, stall))
