"""ctypes binding of include/klsh.h.

`Cluster(rows, min_similarity, cluster_iteration, threads_to_use, dim, bucket_size_threshold,
verbose)` mirrors the reference seam (reference function/cluster.h:42): same argument names and
meaning; `rows` is the (values, id_offsets, ids) triple that stands for `vector<Abundance*>`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

u64 = C.c_uint64
i64 = C.c_int64
f32p = C.POINTER(C.c_float)
u64p = C.POINTER(C.c_uint64)
u16p = C.POINTER(C.c_uint16)
u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)

PLANE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, f32p)
DONE_FN = C.CFUNCTYPE(None, C.c_void_p)


class KlshError(RuntimeError):
    pass


class IterStats(C.Structure):
    _fields_ = [
        ("rows_in", u64),
        ("rows_out", u64),
        ("H", C.c_int32),
        ("threshold", C.c_float),
        ("buckets", u64),
        ("bucket_max", u64),
        ("nested_calls", u64),
        ("eps_margin_rows", u64),
        ("ms_sign", C.c_float),
        ("ms_group", C.c_float),
        ("ms_merge", C.c_float),
        ("ms_compact", C.c_float),
        ("ms_total", C.c_float),
        ("screen_pairs", u64),
        ("exact_pairs", u64),
    ]


class TtestStats(C.Structure):
    _fields_ = [("rows", u64), ("tested", u64), ("rows_a", u64), ("rows_b", u64), ("ids_a", u64), ("ids_b", u64),
                ("margin", u64)]


def lib_path() -> str:
    return os.path.join(_HERE, "libklsh.so")


_LIB = None

# every symbol include/klsh.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("klsh_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("klsh_destroy", None, [C.c_void_p]),
    ("klsh_last_error", C.c_char_p, [C.c_void_p]),
    ("klsh_launch_count", u64, [C.c_void_p]),
    ("klsh_set_seed", C.c_int, [C.c_void_p, u64]),
    ("klsh_set_plane_source", C.c_int, [C.c_void_p, PLANE_FN, C.c_void_p]),
    ("klsh_draw_table", C.c_int, [C.c_void_p, C.c_int, C.c_int, f32p]),
    ("klsh_plane_tell", C.c_int, [C.c_void_p, u64p, u64p]),
    ("klsh_plane_seek", C.c_int, [C.c_void_p, u64, u64]),
    ("klsh_set_draws_done_callback", C.c_int, [C.c_void_p, DONE_FN, C.c_void_p]),
    ("klsh_load_counts", C.c_int, [C.c_void_p, u16p, f32p, C.c_int, u64, u64]),
    ("klsh_set_rows", C.c_int, [C.c_void_p, f32p, u64p, u64p, u64, C.c_int]),
    ("klsh_load_cluster_file", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, u64, u64]),
    ("klsh_cluster", C.c_int, [C.c_void_p, C.c_float, C.c_int, i64, C.POINTER(IterStats)]),
    ("klsh_sign", C.c_int, [C.c_void_p, f32p, u64, C.c_int, f32p, C.c_int, u64p]),
    ("klsh_p_cluster", C.c_int, [C.c_void_p, C.c_float]),
    ("klsh_nested_cluster", C.c_int, [C.c_void_p, C.c_float]),
    ("klsh_cosine_distance", C.c_int, [C.c_void_p, f32p, f32p, u64, C.c_int, f32p]),
    ("klsh_set_consensus", C.c_int, [C.c_void_p, f32p, i64, f32p, i64, C.c_int, f32p]),
    ("klsh_row_count", C.c_int, [C.c_void_p, u64p, u64p]),
    ("klsh_get_rows", C.c_int, [C.c_void_p, f32p, u64p, u64p]),
    ("klsh_save", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, i64]),
    ("klsh_set_id_format", C.c_int, [C.c_void_p, C.c_int]),
    ("klsh_ttest", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, u8p, f64p, f64p, C.POINTER(TtestStats)]),
    ("klsh_differential_ids", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, u64, u8p, C.POINTER(TtestStats)]),
    ("klsh_select_kmers", C.c_int, [C.c_void_p, u8p, u64, C.c_int, u8p, u8p, u64p, u8p, u64p]),
    ("klsh_kmer_set_load", C.c_int, [C.c_void_p, u8p, u64, C.c_int]),
    ("klsh_check_reads", C.c_int, [C.c_void_p, C.c_int, C.c_char_p, u64p, u64, C.c_float, u8p, C.POINTER(C.c_uint32)]),
    ("klsh_stash_rows", C.c_int, [C.c_void_p]),
    ("klsh_stash_count", C.c_int, [C.c_void_p, u64p]),
    ("klsh_unstash_rows", C.c_int, [C.c_void_p]),
    ("klsh_snapshot", C.c_int, [C.c_void_p]),
    ("klsh_restore", C.c_int, [C.c_void_p]),
    ("klsh_sync", C.c_int, [C.c_void_p]),
    ("klsh_row_stride", C.c_int, [C.c_void_p]),
    ("klsh_mg_pass_begin", C.c_int, [C.c_void_p, u64p, C.POINTER(C.c_int32), u64p]),
    ("klsh_mg_plan", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)]),
    ("klsh_mg_merge", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, i64, u64p, u64p, u64p]),
    ("klsh_mg_export", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("klsh_mg_apply", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, u64, C.c_void_p, C.c_void_p, u64]),
    ("klsh_mg_set_alive", C.c_int, [C.c_void_p, C.c_void_p, u64]),
    ("klsh_nccl_unique_id", C.c_int, [C.c_void_p, u64]),
    ("klsh_mg_init", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, u64]),
    ("klsh_mg_finalize", C.c_int, [C.c_void_p]),
    ("klsh_mg_rank", C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("klsh_mg_cluster", C.c_int, [C.c_void_p, C.c_float, C.c_int, i64, C.POINTER(IterStats)]),
    ("klsh_mg_gather_rows", C.c_int, [C.c_void_p]),
]


def load_library():
    """Load libklsh.so.  Raises KlshError if it has not been built — there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise KlshError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(t)


class Context:
    """One GPU context (klsh_ctx)."""

    def __init__(self, device: int = 0, seed: int | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.klsh_create(device, C.byref(h))
        if rc != 0:
            raise KlshError("klsh_create: " + self.lib.klsh_last_error(None).decode())
        self.h = h
        self._cb = None
        self.D = 0
        if seed is not None:
            self.set_seed(seed)

    def _ck(self, rc, what):
        if rc != 0:
            raise KlshError("%s failed (%d): %s" % (what, rc, self.lib.klsh_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.klsh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- planes
    def set_seed(self, seed: int):
        self._ck(self.lib.klsh_set_seed(self.h, seed), "klsh_set_seed")

    def set_plane_source(self, fn):
        """fn(H, D) -> float32 array [H][D]."""

        def tramp(_user, H, D, out):
            t = np.ascontiguousarray(fn(H, D), dtype=np.float32).reshape(H, D)
            C.memmove(out, t.ctypes.data, 4 * H * D)

        self._cb = PLANE_FN(tramp)
        self._ck(self.lib.klsh_set_plane_source(self.h, self._cb, None), "klsh_set_plane_source")

    def plane_tell(self):
        """(seed, hash functions drawn so far) of the built-in hyperplane stream."""
        a, b = u64(), u64()
        self._ck(self.lib.klsh_plane_tell(self.h, C.byref(a), C.byref(b)), "klsh_plane_tell")
        return a.value, b.value

    def plane_seek(self, seed: int, drawn: int):
        self._ck(self.lib.klsh_plane_seek(self.h, seed, drawn), "klsh_plane_seek")

    def set_draws_done_callback(self, fn):
        """fn() is called from inside cluster() once the call has drawn its last hyperplane table."""
        self._done_cb = DONE_FN(lambda _user: fn()) if fn is not None else C.cast(None, DONE_FN)
        self._ck(self.lib.klsh_set_draws_done_callback(self.h, self._done_cb, None), "klsh_set_draws_done_callback")

    def draw_table(self, H: int, D: int):
        out = np.empty((H, D), dtype=np.float32)
        self._ck(self.lib.klsh_draw_table(self.h, H, D, _p(out, f32p)), "klsh_draw_table")
        return out

    # ---- rows in
    def load_counts(self, counts, v_kmers, batch_offset: int = 0):
        counts = np.ascontiguousarray(counts, dtype=np.uint16)
        d, batch = counts.shape
        vk = np.ascontiguousarray(v_kmers, dtype=np.float32)
        assert vk.shape[0] == d
        self.D = d
        self._ck(self.lib.klsh_load_counts(self.h, _p(counts, u16p), _p(vk, f32p), d, batch, batch_offset),
                 "klsh_load_counts")

    def set_rows(self, values, id_offsets=None, ids=None):
        values = np.ascontiguousarray(values, dtype=np.float32)
        n, d = values.shape
        if id_offsets is None:
            id_offsets = np.arange(n + 1, dtype=np.uint64)
            ids = np.arange(n, dtype=np.uint64)
        id_offsets = np.ascontiguousarray(id_offsets, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        self.D = d
        self._ck(self.lib.klsh_set_rows(self.h, _p(values, f32p), _p(id_offsets, u64p), _p(ids, u64p), n, d),
                 "klsh_set_rows")

    def load_cluster_file(self, path: str, d: int, start_line: int = 0, num_lines: int = 0):
        self.D = d
        self._ck(self.lib.klsh_load_cluster_file(self.h, path.encode(), d, start_line, num_lines),
                 "klsh_load_cluster_file")

    # ---- hot path
    def cluster(self, min_similarity: float, iterations: int, bucket_size_threshold: int):
        stats = (IterStats * max(1, iterations))()
        self._ck(self.lib.klsh_cluster(self.h, min_similarity, iterations, bucket_size_threshold, stats), "klsh_cluster")
        return list(stats)

    def sign(self, rows, table):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        table = np.ascontiguousarray(table, dtype=np.float32)
        n, d = rows.shape
        h = table.shape[0]
        keys = np.empty(n, dtype=np.uint64)
        self._ck(self.lib.klsh_sign(self.h, _p(rows, f32p), n, d, _p(table, f32p), h, _p(keys, u64p)), "klsh_sign")
        return keys

    def p_cluster(self, threshold: float):
        self._ck(self.lib.klsh_p_cluster(self.h, threshold), "klsh_p_cluster")

    def nested_cluster(self, threshold: float):
        self._ck(self.lib.klsh_nested_cluster(self.h, threshold), "klsh_nested_cluster")

    def cosine_distance(self, left, right):
        """Distance::cosine for n pairs of rows: 1 - cos, float32 [n]."""
        left = np.ascontiguousarray(left, dtype=np.float32)
        right = np.ascontiguousarray(right, dtype=np.float32)
        n, d = left.shape
        out = np.empty(n, dtype=np.float32)
        self._ck(self.lib.klsh_cosine_distance(self.h, _p(left, f32p), _p(right, f32p), n, d, _p(out, f32p)), "klsh_cosine_distance")
        return out

    def set_consensus(self, current, n_current: int, candidate, n_candidate: int):
        """The centroid AB::SetConsensus(current, candidate) computes for the given member counts."""
        current = np.ascontiguousarray(current, dtype=np.float32)
        candidate = np.ascontiguousarray(candidate, dtype=np.float32)
        out = np.empty_like(current)
        self._ck(self.lib.klsh_set_consensus(self.h, _p(current, f32p), n_current, _p(candidate, f32p), n_candidate, current.shape[0],
                                             _p(out, f32p)), "klsh_set_consensus")
        return out

    # ---- rows out
    def row_count(self, with_ids: bool = True):
        n, m = u64(), u64()
        self._ck(self.lib.klsh_row_count(self.h, C.byref(n), C.byref(m) if with_ids else None), "klsh_row_count")
        return n.value, m.value

    def get_rows(self, out=None):
        """Clusters out: (values[n][D], id_offsets[n+1], ids).  `out` = (values, id_offsets, ids) caller-owned
        host arrays (e.g. pinned, reused across calls) large enough for the result; views of them are returned."""
        n, m = self.row_count()
        if out is None:
            values = np.empty((n, self.D), dtype=np.float32)
            offs = np.empty(n + 1, dtype=np.uint64)
            ids = np.empty(max(m, 1), dtype=np.uint64)
        else:
            bv, bo, bi = out
            if bv.size < n * self.D or bo.size < n + 1 or bi.size < max(m, 1):
                raise KlshError("get_rows: output buffers too small (%d rows, %d ids)" % (n, m))
            values = bv.reshape(-1)[: n * self.D].reshape(n, self.D)
            offs = bo.reshape(-1)[: n + 1]
            ids = bi.reshape(-1)[: max(m, 1)]
        self._ck(self.lib.klsh_get_rows(self.h, _p(values, f32p), _p(offs, u64p), _p(ids, u64p)), "klsh_get_rows")
        return values, offs, ids[:m]

    def save(self, path: str, delfile: bool = True, ignore_small: int = 0):
        self._ck(self.lib.klsh_save(self.h, path.encode(), int(delfile), ignore_small), "klsh_save")

    def set_id_format(self, fmt: int):
        """0: text <F>.clust (the reference's), 1: binary <F>.clust.bin — for save() and load_cluster_file()."""
        self._ck(self.lib.klsh_set_id_format(self.h, fmt), "klsh_set_id_format")

    # ---- mode E statistics (reference app/kmerLSH.cc:541-585, function/funcAB.cc:73-109)
    def ttest(self, num_sample1: int, num_sample2: int, pvalue_thresh: float, size_thresh: int, tails: bool = True):
        """AB::WRS on every cluster of the row set: (row_group, lefttail, righttail, stats)."""
        n, _ = self.row_count(False)
        group = np.zeros(max(n, 1), dtype=np.uint8)
        left = np.empty(max(n, 1), dtype=np.float64) if tails else None
        right = np.empty(max(n, 1), dtype=np.float64) if tails else None
        st = TtestStats()
        self._ck(self.lib.klsh_ttest(self.h, num_sample1, num_sample2, pvalue_thresh, size_thresh, _p(group, u8p),
                                     _p(left, f64p) if tails else None, _p(right, f64p) if tails else None, C.byref(st)),
                 "klsh_ttest")
        return group[:n], (left[:n] if tails else None), (right[:n] if tails else None), st

    def differential_ids(self, num_sample1: int, num_sample2: int, pvalue_thresh: float, size_thresh: int, n_kmers: int):
        """One label per k-mer id < n_kmers: 1 = group-A set, 2 = group-B set, 0 = neither; plus the stats."""
        label = np.zeros(max(n_kmers, 1), dtype=np.uint8)
        st = TtestStats()
        self._ck(self.lib.klsh_differential_ids(self.h, num_sample1, num_sample2, pvalue_thresh, size_thresh, n_kmers,
                                                _p(label, u8p), C.byref(st)), "klsh_differential_ids")
        return label[:n_kmers], st

    def select_kmers(self, records, id_label):
        """The join over kmer_set.hex: (records labelled 1, records labelled 2), each in id order."""
        records = np.ascontiguousarray(records, dtype=np.uint8)
        n, rb = records.shape
        id_label = np.ascontiguousarray(id_label, dtype=np.uint8)
        if id_label.shape[0] != n:
            raise KlshError("select_kmers: %d records but %d labels" % (n, id_label.shape[0]))
        a = np.empty((max(n, 1), rb), dtype=np.uint8)
        b = np.empty((max(n, 1), rb), dtype=np.uint8)
        na, nb = u64(), u64()
        self._ck(self.lib.klsh_select_kmers(self.h, _p(records, u8p), n, rb, _p(id_label, u8p), _p(a, u8p), C.byref(na),
                                            _p(b, u8p), C.byref(nb)), "klsh_select_kmers")
        return a[: na.value].copy(), b[: nb.value].copy()

    # ---- read extraction votes (reference io/ioFastQ.cc:5-75)
    def kmer_set_load(self, records):
        records = np.ascontiguousarray(records, dtype=np.uint8)
        if records.ndim != 2:
            raise KlshError("kmer_set_load: records must be [n][record_bytes]")
        self._ck(self.lib.klsh_kmer_set_load(self.h, _p(records, u8p), records.shape[0], records.shape[1]), "klsh_kmer_set_load")

    def check_reads(self, k: int, seq: bytes, offsets, kmer_vote: float):
        """IOFQ::CheckRead over reads seq[offsets[r]:offsets[r+1]]: (record flags, votes)."""
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offs.shape[0] - 1
        rec = np.zeros(max(n, 1), dtype=np.uint8)
        votes = np.zeros(max(n, 1), dtype=np.uint32)
        self._ck(self.lib.klsh_check_reads(self.h, k, seq, _p(offs, u64p), n, kmer_vote, _p(rec, u8p),
                                           votes.ctypes.data_as(C.POINTER(C.c_uint32))), "klsh_check_reads")
        return rec[:n], votes[:n]

    # ---- survivors of several batches, resident on the device
    def stash_rows(self):
        self._ck(self.lib.klsh_stash_rows(self.h), "klsh_stash_rows")

    def stash_count(self) -> int:
        n = u64()
        self._ck(self.lib.klsh_stash_count(self.h, C.byref(n)), "klsh_stash_count")
        return n.value

    def unstash_rows(self):
        self._ck(self.lib.klsh_unstash_rows(self.h), "klsh_unstash_rows")

    # ---- state
    def snapshot(self):
        self._ck(self.lib.klsh_snapshot(self.h), "klsh_snapshot")

    def restore(self):
        self._ck(self.lib.klsh_restore(self.h), "klsh_restore")

    def sync(self):
        self._ck(self.lib.klsh_sync(self.h), "klsh_sync")

    def launch_count(self) -> int:
        return self.lib.klsh_launch_count(self.h)

    # ---- multi-GPU building blocks (device pointers are passed as integers, e.g. tensor.data_ptr())
    def row_stride(self) -> int:
        return self.lib.klsh_row_stride(self.h)

    def mg_pass_begin(self):
        n, H, nb = u64(), C.c_int32(), u64()
        self._ck(self.lib.klsh_mg_pass_begin(self.h, C.byref(n), C.byref(H), C.byref(nb)), "klsh_mg_pass_begin")
        return n.value, H.value, nb.value

    def mg_plan(self, world: int):
        out = (C.c_uint32 * (world + 1))()
        self._ck(self.lib.klsh_mg_plan(self.h, world, out), "klsh_mg_plan")
        return list(out)

    def mg_merge(self, b_lo: int, b_hi: int, threshold: float, bucket_size_threshold: int):
        a, b, c = u64(), u64(), u64()
        self._ck(self.lib.klsh_mg_merge(self.h, b_lo, b_hi, threshold, bucket_size_threshold, C.byref(a), C.byref(b), C.byref(c)),
                 "klsh_mg_merge")
        return a.value, b.value, c.value

    def mg_export(self, d_surv, d_mod_rows, d_mod_vals, d_mod_meta, d_slots, d_vals):
        self._ck(self.lib.klsh_mg_export(self.h, d_surv, d_mod_rows, d_mod_vals, d_mod_meta, d_slots, d_vals), "klsh_mg_export")

    def mg_apply(self, d_mod_rows, d_mod_vals, d_mod_meta, n_mod, d_slots, d_vals, n_next):
        self._ck(self.lib.klsh_mg_apply(self.h, d_mod_rows, d_mod_vals, d_mod_meta, n_mod, d_slots, d_vals, n_next), "klsh_mg_apply")

    def mg_set_alive(self, d_alive, n: int):
        self._ck(self.lib.klsh_mg_set_alive(self.h, d_alive, n), "klsh_mg_set_alive")

    # ---- multi-GPU with NCCL inside the library
    def mg_init(self, rank: int, world: int, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.lib.klsh_mg_init(self.h, rank, world, buf, 128), "klsh_mg_init")

    def mg_finalize(self):
        self._ck(self.lib.klsh_mg_finalize(self.h), "klsh_mg_finalize")

    def mg_cluster(self, min_similarity: float, iterations: int, bucket_size_threshold: int):
        stats = (IterStats * max(1, iterations))()
        self._ck(self.lib.klsh_mg_cluster(self.h, min_similarity, iterations, bucket_size_threshold, stats), "klsh_mg_cluster")
        return list(stats)

    def mg_gather_rows(self):
        self._ck(self.lib.klsh_mg_gather_rows(self.h), "klsh_mg_gather_rows")


def nccl_unique_id() -> bytes:
    """128-byte NCCL unique id (create on rank 0, hand to every rank)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.klsh_nccl_unique_id(buf, 128)
    if rc != 0:
        raise KlshError("klsh_nccl_unique_id: " + lib.klsh_last_error(None).decode())
    return buf.raw


def Cluster(rows, min_similarity, cluster_iteration, threads_to_use, dim, bucket_size_threshold, verbose=False,
            seed=None, plane_source=None, device=0):
    """Drop-in for the reference's Cluster() (function/cluster.cc:181-340).

    rows = (values[n][dim], id_offsets[n+1], ids) in, the clustered triple out (the reference
    mutates its vector in place).  threads_to_use is accepted for signature compatibility; results
    equal the reference's single-thread schedule.  Hyperplanes come from `seed` (the reference
    generator behind a seeded random_device) or from `plane_source(H, D)`.
    """
    values, id_offsets, ids = rows
    values = np.asarray(values, dtype=np.float32)
    assert values.shape[1] == dim
    with Context(device) as ctx:
        if plane_source is not None:
            ctx.set_plane_source(plane_source)
        elif seed is not None:
            ctx.set_seed(seed)
        ctx.set_rows(values, id_offsets, ids)
        stats = ctx.cluster(min_similarity, cluster_iteration, bucket_size_threshold)
        if verbose:
            for k, s in enumerate(stats):
                print("Iteration:\t%d, cos sim threshold:\t%g dimension : %d" % (k + 1, s.threshold, dim))
                print("Size of profilings : %d" % s.rows_in)
                print("#k-mers after clustering:\t%d" % s.rows_out)
        return ctx.get_rows()
