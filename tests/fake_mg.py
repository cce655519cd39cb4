"""CPU test double of the multi-GPU building blocks (klsh_mg_*), built from oracle primitives.
It lets the sharded-Cluster protocol in kmerlsh_b200/distributed.py run under gloo without a GPU:
same methods, same exchanged arrays (survivors, modified rows, member-chain writes), numpy state."""
import numpy as np
import torch


class FakeMgBackend:
    def __init__(self, oracle, values, seed):
        self.o = oracle
        self.vals = np.ascontiguousarray(values, dtype=np.float32).copy()
        n, self.D = self.vals.shape
        self.cnt = np.ones(n, np.int32)
        self.head = np.arange(n, dtype=np.int32)
        self.tail = np.arange(n, dtype=np.int32)
        self.next = np.full(n, -1, np.int32)
        self.alive = np.arange(n, dtype=np.uint32)
        self.planes = oracle.planes(seed)

    # ---- protocol methods -------------------------------------------------------------------
    def mg_pass_begin(self):
        n = len(self.alive)
        if n == 0:
            return 0, 0, 0
        H = n.bit_length() - 1
        table = self.planes.table(H, self.D)
        keys = self.o.sign(self.vals[self.alive], table) if H else np.zeros(n, np.uint32)
        order = np.argsort(keys, kind="stable")
        self.rows_sorted = self.alive[order].copy()
        ks = keys[order]
        self.bstart = np.concatenate([[0], np.flatnonzero(ks[1:] != ks[:-1]) + 1, [n]]).astype(np.int64)
        self.n = n
        return n, H, len(self.bstart) - 1

    def mg_plan(self, world):
        nb = len(self.bstart) - 1
        splits = [int(np.searchsorted(self.bstart[:nb], r * self.n // world, side="left")) for r in range(world)]
        splits[0] = 0
        return splits + [nb]

    def _p_cluster(self, rows, thr):
        """reference p_cluster (function/cluster.cc:56-87) on a list of row indices, logging changes"""
        c = list(rows)
        size, i = len(c), 1
        thr = np.float32(thr)
        while i < size:
            cur = c[i]
            merged = False
            for j in range(i):
                cand = c[j]
                d = self.o.cosine_distance(self.vals[cur], self.vals[cand])
                if np.float32(1) - d >= thr:
                    self.vals[cand] = self.o.consensus(self.vals[cur], int(self.cnt[cur]), self.vals[cand], int(self.cnt[cand]))
                    t1 = int(self.tail[cur])
                    self.next[t1] = self.head[cand]
                    self.chain_log.append((t1, int(self.head[cand])))
                    self.head[cand] = self.head[cur]
                    self.cnt[cand] += self.cnt[cur]
                    self.mod_log.append(cand)
                    size -= 1
                    c[i] = c[size]
                    merged = True
                    break
            if not merged:
                i += 1
        return c[:size]

    def _nested(self, rows, thr, table, H2):
        keys = self.o.sign(self.vals[rows], table) if H2 else np.zeros(len(rows), np.uint32)
        order = np.argsort(keys, kind="stable")
        rows = np.asarray(rows)[order]
        ks = keys[order]
        bs = np.concatenate([[0], np.flatnonzero(ks[1:] != ks[:-1]) + 1, [len(rows)]])
        out = []
        for b in range(len(bs) - 1):
            out += self._p_cluster(rows[bs[b]:bs[b + 1]], thr)
        return out

    def mg_merge(self, b_lo, b_hi, thr, nest):
        self.mod_log, self.chain_log = [], []
        nb = len(self.bstart) - 1
        sizes = np.diff(self.bstart)
        results = {}
        # every rank draws the table of every oversized bucket, in order
        if nest >= 0:
            for b in np.flatnonzero(sizes > max(nest, 1)):
                H2 = int(sizes[b]).bit_length() - 1
                table = self.planes.table(H2, self.D)
                if b_lo <= b < b_hi:
                    results[int(b)] = self._nested(self.rows_sorted[self.bstart[b]:self.bstart[b + 1]], thr, table, H2)
        surv = []
        for b in range(b_lo, min(b_hi, nb)):
            rows = self.rows_sorted[self.bstart[b]:self.bstart[b + 1]]
            if b in results:
                surv += results[b]
            elif len(rows) >= 2:
                surv += self._p_cluster(rows, thr)
            else:
                surv += list(rows)
        self._surv = np.asarray(surv, dtype=np.uint32)
        return len(surv), len(self.mod_log), len(self.chain_log)

    def mg_export(self, n_surv, n_mod, n_chain):
        mod = np.asarray(self.mod_log, dtype=np.int64)
        meta = np.stack([self.cnt[mod], self.head[mod], self.tail[mod]], axis=1).astype(np.int32) if n_mod else np.zeros((0, 3), np.int32)
        chain = np.asarray(self.chain_log, dtype=np.int64).reshape(-1, 2)
        t = torch.from_numpy
        return [torch.tensor([n_surv, n_mod, n_chain], dtype=torch.int64), t(self._surv.astype(np.int32)),
                t(mod.astype(np.int32)), t(self.vals[mod].copy() if n_mod else np.zeros((0, self.D), np.float32)), t(meta),
                t(chain[:, 0].astype(np.int32)), t(chain[:, 1].astype(np.int32))]

    def to_host(self, counts):
        return counts.tolist()

    def mg_apply(self, mod_rows, mod_vals, mod_meta, nm, slots, vals, nc):
        r = mod_rows.numpy()[:nm].astype(np.int64)
        self.vals[r] = mod_vals.numpy()[:nm]
        m = mod_meta.numpy()[:nm]
        self.cnt[r], self.head[r], self.tail[r] = m[:, 0], m[:, 1], m[:, 2]
        self.next[slots.numpy()[:nc].astype(np.int64)] = vals.numpy()[:nc]

    def mg_set_alive(self, parts, total):
        self.alive = np.concatenate([s.numpy()[:n].astype(np.uint32) for s, n in parts]) if total else np.zeros(0, np.uint32)

    # ---- result -----------------------------------------------------------------------------
    def get_rows(self):
        offs, ids = [0], []
        for r in self.alive:
            s = int(self.head[r])
            while s >= 0:
                ids.append(s)
                s = int(self.next[s])
            offs.append(len(ids))
        return self.vals[self.alive].copy(), np.asarray(offs, np.uint64), np.asarray(ids, np.uint64)
