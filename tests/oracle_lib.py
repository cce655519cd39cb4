"""ctypes access to the checkers (TEST INFRASTRUCTURE):

* `Oracle`  -> oracle/libklsh_oracle.so, the C restatement (oracle/klsh_oracle.c)
* `RefLib`  -> oracle/_ref/libklsh_ref.so, the reference's own objects behind oracle/ref_harness.cc
               (present only where /root/reference was available to build it)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libklsh_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libklsh_ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "kmerLSH_ref")

u64 = C.c_uint64
i64 = C.c_int64
f32p = C.POINTER(C.c_float)
u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u16p = C.POINTER(C.c_uint16)


def _p(a, t):
    return a.ctypes.data_as(t)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(ORACLE_DIR, "klsh_oracle.c")
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


class IterStats(C.Structure):
    _fields_ = [
        ("rows_in", u64),
        ("rows_out", u64),
        ("H", C.c_int),
        ("threshold", C.c_float),
        ("buckets_nonempty", u64),
        ("bucket_max", u64),
        ("nested_calls", u64),
        ("compares", u64),
        ("merges", u64),
    ]


def flat_ids(n):
    """id_offsets/ids for n singleton rows with ids 0..n-1."""
    return np.arange(n + 1, dtype=np.uint64), np.arange(n, dtype=np.uint64)


class Oracle:
    def __init__(self):
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        self.L = L
        L.klo_planes_new.restype = C.c_void_p
        L.klo_planes_new.argtypes = [u64]
        L.klo_planes_free.argtypes = [C.c_void_p]
        L.klo_planes_reseed.argtypes = [C.c_void_p, u64]
        L.klo_planes_draws.restype = u64
        L.klo_planes_draws.argtypes = [C.c_void_p]
        L.klo_planes_table.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p]
        L.klo_log_lut.argtypes = [f32p]
        L.klo_vkmers.argtypes = [f32p, u64, C.c_int, f32p]
        L.klo_convert_counts.restype = u64
        L.klo_convert_counts.argtypes = [u16p, f32p, C.c_int, u64, u64, f32p, u64p]
        L.klo_sign.argtypes = [f32p, u64, C.c_int, f32p, C.c_int, u32p]
        L.klo_cosine_distance.restype = C.c_float
        L.klo_cosine_distance.argtypes = [f32p, f32p, C.c_int]
        L.klo_consensus.argtypes = [f32p, i64, f32p, i64, C.c_int, f32p]
        L.klo_threshold_after.restype = C.c_float
        L.klo_threshold_after.argtypes = [C.c_float, C.c_int, C.c_int]
        L.klo_rows_new.restype = C.c_void_p
        L.klo_rows_new.argtypes = [f32p, u64p, u64p, u64, C.c_int]
        L.klo_rows_free.argtypes = [C.c_void_p]
        L.klo_rows_count.restype = u64
        L.klo_rows_count.argtypes = [C.c_void_p]
        L.klo_rows_members.restype = u64
        L.klo_rows_members.argtypes = [C.c_void_p]
        L.klo_rows_export.argtypes = [C.c_void_p, f32p, u64p, u64p]
        L.klo_p_cluster.argtypes = [C.c_void_p, C.c_float]
        L.klo_nested_cluster.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.klo_cluster.argtypes = [C.c_void_p, C.c_float, C.c_int, i64, C.c_void_p, C.POINTER(IterStats)]
        L.klo_bucket_sizes.restype = u64
        L.klo_bucket_sizes.argtypes = [C.c_void_p, f32p, C.c_int, u64p, u64]
        L.klo_save.restype = C.c_int
        L.klo_save.argtypes = [C.c_void_p, C.c_char_p, C.c_int, i64]
        L.klo_read_cluster.restype = C.c_void_p
        L.klo_read_cluster.argtypes = [C.c_char_p, C.c_int, u64, u64]
        f64p = C.POINTER(C.c_double)
        u8p = C.POINTER(C.c_uint8)
        L.klo_ttest2.argtypes = [f64p, C.c_int, f64p, C.c_int, f64p, f64p, f64p]
        L.klo_wrs_rows.argtypes = [f32p, u64p, u64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, u8p, f64p, f64p]
        L.klo_differential_ids.argtypes = [u8p, u64p, u64p, u64, u64, u8p]
        L.klo_select_kmers.argtypes = [u8p, u64, C.c_int, u8p, u8p, u64p, u8p, u64p]
        L.klo_kmer_from_string.argtypes = [C.c_char_p, C.c_int, u8p]
        L.klo_kmer_rep.argtypes = [u8p, C.c_int, u8p]
        L.klo_check_reads.argtypes = [u8p, u64, C.c_int, C.c_char_p, u64p, u64, C.c_float, u8p, u32p]
        L.klo_mode_c.restype = C.c_int
        L.klo_mode_c.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_float, C.c_int, C.c_char_p, C.c_char_p,
                                 u64, i64, u64, C.POINTER(IterStats)]

    # ---- planes
    def planes(self, seed):
        return _Planes(self, seed)

    # ---- transform
    def log_lut(self):
        lut = np.empty(65536, dtype=np.float32)
        self.L.klo_log_lut(_p(lut, f32p))
        return lut

    def convert_counts(self, counts, v_kmers, batch_offset=0):
        counts = np.ascontiguousarray(counts, dtype=np.uint16)
        d, batch = counts.shape
        vk = np.ascontiguousarray(v_kmers, dtype=np.float32)
        values = np.empty((batch, d), dtype=np.float32)
        ids = np.empty(batch, dtype=np.uint64)
        k = self.L.klo_convert_counts(_p(counts, u16p), _p(vk, f32p), d, batch, batch_offset, _p(values, f32p),
                                      _p(ids, u64p))
        return values[:k].copy(), ids[:k].copy()

    # ---- scalar kernels
    def sign(self, rows, table):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        table = np.ascontiguousarray(table, dtype=np.float32)
        n, d = rows.shape
        h = table.shape[0]
        keys = np.empty(n, dtype=np.uint32)
        self.L.klo_sign(_p(rows, f32p), n, d, _p(table, f32p), h, _p(keys, u32p))
        return keys

    def cosine_distance(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        return np.float32(self.L.klo_cosine_distance(_p(a, f32p), _p(b, f32p), a.shape[0]))

    def consensus(self, cur, c1, cand, c2):
        cur = np.ascontiguousarray(cur, dtype=np.float32)
        cand = np.ascontiguousarray(cand, dtype=np.float32)
        out = np.empty_like(cur)
        self.L.klo_consensus(_p(cur, f32p), c1, _p(cand, f32p), c2, cur.shape[0], _p(out, f32p))
        return out

    def threshold_after(self, min_similarity, iterations, steps):
        return np.float32(self.L.klo_threshold_after(min_similarity, iterations, steps))

    # ---- mode E statistics
    def ttest2(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        out = np.zeros(3, dtype=np.float64)
        f64p = C.POINTER(C.c_double)
        self.L.klo_ttest2(_p(x, f64p), x.shape[0], _p(y, f64p), y.shape[0], _p(out[0:], f64p), _p(out[1:], f64p), _p(out[2:], f64p))
        return float(out[0]), float(out[1]), float(out[2])

    def wrs_rows(self, values, id_offsets, n1, n2, pvalue_thresh, size_thresh):
        values = np.ascontiguousarray(values, dtype=np.float32)
        offs = np.ascontiguousarray(id_offsets, dtype=np.uint64)
        n, d = values.shape
        group = np.zeros(n, dtype=np.uint8)
        left = np.empty(n, dtype=np.float64)
        right = np.empty(n, dtype=np.float64)
        f64p = C.POINTER(C.c_double)
        self.L.klo_wrs_rows(_p(values, f32p), _p(offs, u64p), n, d, n1, n2, pvalue_thresh, size_thresh,
                            _p(group, C.POINTER(C.c_uint8)), _p(left, f64p), _p(right, f64p))
        return group, left, right

    def differential_ids(self, group, id_offsets, ids, n_kmers):
        group = np.ascontiguousarray(group, dtype=np.uint8)
        offs = np.ascontiguousarray(id_offsets, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        label = np.zeros(max(n_kmers, 1), dtype=np.uint8)
        u8p = C.POINTER(C.c_uint8)
        self.L.klo_differential_ids(_p(group, u8p), _p(offs, u64p), _p(ids, u64p), group.shape[0], n_kmers, _p(label, u8p))
        return label[:n_kmers]

    def select_kmers(self, records, id_label):
        records = np.ascontiguousarray(records, dtype=np.uint8)
        n, rb = records.shape
        id_label = np.ascontiguousarray(id_label, dtype=np.uint8)
        a = np.empty((max(n, 1), rb), dtype=np.uint8)
        b = np.empty((max(n, 1), rb), dtype=np.uint8)
        na, nb = u64(0), u64(0)
        u8p = C.POINTER(C.c_uint8)
        self.L.klo_select_kmers(_p(records, u8p), n, rb, _p(id_label, u8p), _p(a, u8p), C.byref(na), _p(b, u8p), C.byref(nb))
        return a[: na.value].copy(), b[: nb.value].copy()

    # ---- read extraction votes
    def kmer_rep(self, s: bytes, k: int):
        """(Kmer(s) bytes, canonical rep bytes), 8 bytes each."""
        km = np.zeros(8, dtype=np.uint8)
        rep = np.zeros(8, dtype=np.uint8)
        u8p = C.POINTER(C.c_uint8)
        self.L.klo_kmer_from_string(s, k, _p(km, u8p))
        self.L.klo_kmer_rep(_p(km, u8p), k, _p(rep, u8p))
        return km, rep

    def check_reads(self, kmers, k, seq: bytes, offsets, kmer_vote):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint8).reshape(-1, 8)
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offs.shape[0] - 1
        rec = np.zeros(max(n, 1), dtype=np.uint8)
        votes = np.zeros(max(n, 1), dtype=np.uint32)
        self.L.klo_check_reads(_p(kmers, C.POINTER(C.c_uint8)), kmers.shape[0], k, seq, _p(offs, u64p), n, kmer_vote,
                               _p(rec, C.POINTER(C.c_uint8)), _p(votes, u32p))
        return rec[:n], votes[:n]

    # ---- row sets
    def rows(self, values, id_offsets=None, ids=None):
        return _Rows(self, values, id_offsets, ids)

    def read_cluster(self, path, d, start=0, num=0):
        h = self.L.klo_read_cluster(path.encode(), d, start, num)
        if not h:
            raise IOError(path)
        return _Rows(self, None, None, None, handle=h, d=d)

    def mode_c(self, work_dir, d, min_similarity, iterations, out_path, seed, batch_thresh=100_000_000,
               phase2_bucket_threshold=1_000_000, tmp_dir=None):
        tmp_dir = tmp_dir or os.path.join(work_dir, "tmp") + "/"
        stats = (IterStats * max(1, iterations))()
        rc = self.L.klo_mode_c(os.path.join(work_dir, "kmer_count.bin").encode(),
                               os.path.join(work_dir, "kmer_count.log").encode(), d, min_similarity, iterations,
                               tmp_dir.encode(), out_path.encode(), batch_thresh, phase2_bucket_threshold, seed,
                               stats)
        if rc != 0:
            raise RuntimeError("klo_mode_c rc=%d" % rc)
        return list(stats)


class _Planes:
    def __init__(self, o, seed):
        self.o = o
        self.h = o.L.klo_planes_new(seed)

    def table(self, H, D):
        out = np.empty((H, D), dtype=np.float32)
        self.o.L.klo_planes_table(self.h, H, D, _p(out, f32p))
        return out

    def draws(self):
        return self.o.L.klo_planes_draws(self.h)

    def __del__(self):
        if self.h:
            self.o.L.klo_planes_free(self.h)
            self.h = None


class _Rows:
    def __init__(self, o, values, id_offsets, ids, handle=None, d=None):
        self.o = o
        if handle is not None:
            self.h, self.d = handle, d
            return
        values = np.ascontiguousarray(values, dtype=np.float32)
        n, d = values.shape
        if id_offsets is None:
            id_offsets, ids = flat_ids(n)
        id_offsets = np.ascontiguousarray(id_offsets, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        self.d = d
        self.h = o.L.klo_rows_new(_p(values, f32p), _p(id_offsets, u64p), _p(ids, u64p), n, d)

    def __len__(self):
        return self.o.L.klo_rows_count(self.h)

    def export(self):
        n = len(self)
        m = self.o.L.klo_rows_members(self.h)
        values = np.empty((n, self.d), dtype=np.float32)
        offs = np.empty(n + 1, dtype=np.uint64)
        ids = np.empty(max(m, 1), dtype=np.uint64)
        self.o.L.klo_rows_export(self.h, _p(values, f32p), _p(offs, u64p), _p(ids, u64p))
        return values, offs, ids[:m]

    def p_cluster(self, threshold):
        self.o.L.klo_p_cluster(self.h, threshold)

    def nested_cluster(self, threshold, planes):
        self.o.L.klo_nested_cluster(self.h, threshold, planes.h)

    def cluster(self, min_similarity, iterations, bucket_size_threshold, planes):
        stats = (IterStats * max(1, iterations))()
        self.o.L.klo_cluster(self.h, min_similarity, iterations, bucket_size_threshold, planes.h, stats)
        return list(stats)

    def bucket_sizes(self, table):
        table = np.ascontiguousarray(table, dtype=np.float32)
        n = len(self)
        sizes = np.empty(max(n, 1), dtype=np.uint64)
        nb = self.o.L.klo_bucket_sizes(self.h, _p(table, f32p), table.shape[0], _p(sizes, u64p), n)
        return sizes[:nb]

    def save(self, path, delfile=True, ignore_small=0):
        rc = self.o.L.klo_save(self.h, path.encode(), int(delfile), ignore_small)
        if rc != 0:
            raise IOError(path)

    def __del__(self):
        if getattr(self, "h", None):
            self.o.L.klo_rows_free(self.h)
            self.h = None


class RefLib:
    """The reference's own functions (oracle/ref_harness.cc).  Must be created in a process whose
    environment had OMP_THREAD_LIMIT=1 before libgomp was loaded."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        os.environ.setdefault("OMP_THREAD_LIMIT", "1")
        L = C.CDLL(REF_SO)
        self.L = L
        L.ref_reseed.argtypes = [C.c_ulonglong]
        L.ref_master_draws.restype = C.c_ulonglong
        L.ref_generate_table.argtypes = [C.c_int, C.c_int, f32p]
        L.ref_random_projection.argtypes = [f32p, u64, C.c_int, f32p, C.c_int, C.POINTER(C.c_int)]
        L.ref_cosine.restype = C.c_float
        L.ref_cosine.argtypes = [f32p, f32p, C.c_int]
        L.ref_set_consensus.argtypes = [f32p, u64, f32p, u64, C.c_int, f32p]
        L.ref_p_cluster.argtypes = [f32p, u64p, u64p, u64, C.c_int, C.c_float]
        L.ref_nested_cluster.argtypes = [f32p, u64p, u64p, u64, C.c_int, C.c_float, C.c_int]
        L.ref_cluster.argtypes = [f32p, u64p, u64p, u64, C.c_int, C.c_float, C.c_int, C.c_uint, C.c_int, C.c_int]
        L.ref_convert_ht_mat.argtypes = [u16p, f32p, C.c_int, u64, u64]
        L.ref_read_cluster_all.argtypes = [C.c_char_p, C.c_int]
        L.ref_result_rows.restype = u64
        L.ref_result_ids.restype = u64
        L.ref_result_copy.argtypes = [f32p, u64p, u64p, C.c_int]
        L.ref_save.argtypes = [C.c_char_p, C.c_int, C.c_int]
        f64p = C.POINTER(C.c_double)
        u8p = C.POINTER(C.c_uint8)
        if hasattr(L, "ref_wrs"):
            L.ref_studentttest2.argtypes = [f64p, C.c_int, f64p, C.c_int, f64p, f64p, f64p]
            L.ref_wrs.argtypes = [f32p, u64p, u64p, u64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, u64, u8p, u8p]

    def kmer_rep(self, s: bytes, k: int):
        """Kmer(s) and rep = min(km, twin) as 8-byte records; k is fixed by the first call in the process."""
        self.L.ref_kmer_rep.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
        km = np.zeros(8, dtype=np.uint8)
        rep = np.zeros(8, dtype=np.uint8)
        if self.L.ref_kmer_rep(s, k, _p(km, C.POINTER(C.c_uint8)), _p(rep, C.POINTER(C.c_uint8))) != 0:
            raise RuntimeError("Kmer::k is already set to another value in this process")
        return km, rep

    def check_reads(self, kmers, k, seq: bytes, offsets, kmer_vote):
        self.L.ref_check_reads.argtypes = [C.POINTER(C.c_uint8), u64, C.c_int, C.c_char_p, u64p, u64, C.c_float, C.POINTER(C.c_uint8)]
        kmers = np.ascontiguousarray(kmers, dtype=np.uint8).reshape(-1, 8)
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offs.shape[0] - 1
        rec = np.zeros(max(n, 1), dtype=np.uint8)
        rc = self.L.ref_check_reads(_p(kmers, C.POINTER(C.c_uint8)), kmers.shape[0], k, seq, _p(offs, u64p), n, kmer_vote,
                                    _p(rec, C.POINTER(C.c_uint8)))
        if rc != 0:
            raise RuntimeError("ref_check_reads rc=%d" % rc)
        return rec[:n]

    def ttest2(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        out = np.zeros(3, dtype=np.float64)
        f64p = C.POINTER(C.c_double)
        self.L.ref_studentttest2(_p(x, f64p), x.shape[0], _p(y, f64p), y.shape[0], _p(out[0:], f64p), _p(out[1:], f64p), _p(out[2:], f64p))
        return float(out[0]), float(out[1]), float(out[2])

    def wrs(self, values, id_offsets, ids, n1, n2, pvalue_thresh, size_thresh, n_kmers):
        v, o, i = self._in(values, id_offsets, ids)
        group = np.zeros(v.shape[0], dtype=np.uint8)
        label = np.zeros(max(n_kmers, 1), dtype=np.uint8)
        u8p = C.POINTER(C.c_uint8)
        self.L.ref_wrs(_p(v, f32p), _p(o, u64p), _p(i, u64p), v.shape[0], v.shape[1], n1, n2, pvalue_thresh, size_thresh,
                       n_kmers, _p(group, u8p), _p(label, u8p))
        return group, label[:n_kmers]

    def reseed(self, seed):
        self.L.ref_reseed(seed)

    def draws(self):
        return self.L.ref_master_draws()

    def table(self, H, D):
        out = np.empty((H, D), dtype=np.float32)
        self.L.ref_generate_table(H, D, _p(out, f32p))
        return out

    def sign(self, rows, table):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        table = np.ascontiguousarray(table, dtype=np.float32)
        keys = np.empty(rows.shape[0], dtype=np.int32)
        self.L.ref_random_projection(_p(rows, f32p), rows.shape[0], rows.shape[1], _p(table, f32p), table.shape[0],
                                     keys.ctypes.data_as(C.POINTER(C.c_int)))
        return keys.astype(np.uint32)

    def cosine_distance(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        return np.float32(self.L.ref_cosine(_p(a, f32p), _p(b, f32p), a.shape[0]))

    def consensus(self, cur, c1, cand, c2):
        cur = np.ascontiguousarray(cur, dtype=np.float32)
        cand = np.ascontiguousarray(cand, dtype=np.float32)
        out = np.empty_like(cur)
        self.L.ref_set_consensus(_p(cur, f32p), c1, _p(cand, f32p), c2, cur.shape[0], _p(out, f32p))
        return out

    def _result(self, d):
        n = self.L.ref_result_rows()
        m = self.L.ref_result_ids()
        values = np.empty((n, d), dtype=np.float32)
        offs = np.empty(n + 1, dtype=np.uint64)
        ids = np.empty(max(m, 1), dtype=np.uint64)
        self.L.ref_result_copy(_p(values, f32p), _p(offs, u64p), _p(ids, u64p), d)
        return values, offs, ids[:m]

    def _in(self, values, id_offsets, ids):
        values = np.ascontiguousarray(values, dtype=np.float32)
        if id_offsets is None:
            id_offsets, ids = flat_ids(values.shape[0])
        return values, np.ascontiguousarray(id_offsets, dtype=np.uint64), np.ascontiguousarray(ids, dtype=np.uint64)

    def p_cluster(self, values, threshold, id_offsets=None, ids=None):
        v, o, i = self._in(values, id_offsets, ids)
        self.L.ref_p_cluster(_p(v, f32p), _p(o, u64p), _p(i, u64p), v.shape[0], v.shape[1], threshold)
        return self._result(v.shape[1])

    def nested_cluster(self, values, threshold, id_offsets=None, ids=None):
        v, o, i = self._in(values, id_offsets, ids)
        self.L.ref_nested_cluster(_p(v, f32p), _p(o, u64p), _p(i, u64p), v.shape[0], v.shape[1], threshold, 1)
        return self._result(v.shape[1])

    def cluster(self, values, min_similarity, iterations, bucket_size_threshold, id_offsets=None, ids=None,
                threads=1):
        v, o, i = self._in(values, id_offsets, ids)
        self.L.ref_cluster(_p(v, f32p), _p(o, u64p), _p(i, u64p), v.shape[0], v.shape[1], min_similarity,
                           iterations, threads, bucket_size_threshold, 0)
        return self._result(v.shape[1])

    def convert_counts(self, counts, v_kmers, batch_offset=0):
        counts = np.ascontiguousarray(counts, dtype=np.uint16)
        d, batch = counts.shape
        vk = np.ascontiguousarray(v_kmers, dtype=np.float32)
        self.L.ref_convert_ht_mat(_p(counts, u16p), _p(vk, f32p), d, batch, batch_offset)
        v, o, i = self._result(d)
        return v, i

    def save(self, path, delfile=True, ignore_small=0):
        self.L.ref_save(path.encode(), int(delfile), ignore_small)
