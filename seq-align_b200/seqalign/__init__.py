"""seqalign -- Python binding of the B200-native seq-align hot path.

Thin ctypes layer over ``libseqalign_b200.so`` (CUDA engine + C-ABI, see
``include/seqalign_b200.h``) that mirrors the reference's C API names:
``Scoring`` wraps ``scoring_t`` (reference src/alignment_scoring.h:19-40),
``needleman_wunsch`` / ``smith_waterman`` are the single-pair calls
(reference src/needleman_wunsch.h:22-32, src/smith_waterman.h:21-39) and
``BatchAligner`` is the batch entry point the GPU needs.

There is no CPU implementation behind this module: loading fails loudly if
the shared library has not been built (``make``), and creating an engine
fails loudly when no sm_100 device is usable.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFAULT_LIB = os.path.join(os.path.dirname(_HERE), "lib", "libseqalign_b200.so")

NW = 0
SW = 1
MODE_SCORE = 0
MODE_ALIGN = 1
MODE_SCORE_ONLY = 2
MODE_HITS = 3
MODE_MATS = 4

ERR_CUDA = -1
ERR_UNKNOWN_PAIR = -2
ERR_TRACEBACK = -3
ERR_ARG = -4
ERR_NOMEM = -5


class SeqAlignError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("seqalign error %d: %s" % (code, message))
        self.code = code


class ScoringT(ctypes.Structure):
    """Field-for-field image of scoring_t (the layout is ABI)."""

    _fields_ = [
        ("gap_open", ctypes.c_int),
        ("gap_extend", ctypes.c_int),
        ("no_start_gap_penalty", ctypes.c_bool),
        ("no_end_gap_penalty", ctypes.c_bool),
        ("no_gaps_in_a", ctypes.c_bool),
        ("no_gaps_in_b", ctypes.c_bool),
        ("no_mismatches", ctypes.c_bool),
        ("use_match_mismatch", ctypes.c_bool),
        ("match", ctypes.c_int),
        ("mismatch", ctypes.c_int),
        ("case_sensitive", ctypes.c_bool),
        ("wildcards", ctypes.c_uint32 * 8),
        ("swap_set", (ctypes.c_uint32 * 8) * 256),
        ("wildscores", ctypes.c_int * 256),
        ("swap_scores", (ctypes.c_int * 256) * 256),
        ("min_penalty", ctypes.c_int),
        ("max_penalty", ctypes.c_int),
    ]


class AlignmentT(ctypes.Structure):
    """alignment_t (reference src/alignment.h:33-40)."""

    _fields_ = [
        ("result_a", ctypes.c_void_p),
        ("result_b", ctypes.c_void_p),
        ("capacity", ctypes.c_size_t),
        ("length", ctypes.c_size_t),
        ("pos_a", ctypes.c_size_t),
        ("pos_b", ctypes.c_size_t),
        ("len_a", ctypes.c_size_t),
        ("len_b", ctypes.c_size_t),
        ("score", ctypes.c_int),
    ]


class AlignerT(ctypes.Structure):
    """aligner_t (reference src/alignment.h:23-30)."""

    _fields_ = [
        ("scoring", ctypes.c_void_p),
        ("seq_a", ctypes.c_void_p),
        ("seq_b", ctypes.c_void_p),
        ("score_width", ctypes.c_size_t),
        ("score_height", ctypes.c_size_t),
        ("match_scores", ctypes.POINTER(ctypes.c_int)),
        ("gap_a_scores", ctypes.POINTER(ctypes.c_int)),
        ("gap_b_scores", ctypes.POINTER(ctypes.c_int)),
        ("capacity", ctypes.c_size_t),
    ]


_lib = None


def lib_path():
    return os.environ.get("SEQALIGN_LIB", _DEFAULT_LIB)


def load():
    """Load the shared library (once).  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            "%s not found: build it with `make` (nvcc, sm_100a). "
            "This package has no CPU implementation." % path
        )
    L = ctypes.CDLL(path)
    vp, sz, i32p = ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int32)
    L.seqalign_device_count.restype = ctypes.c_int
    L.seqalign_version.restype = ctypes.c_char_p
    L.seqalign_enable_peer_access.argtypes = [ctypes.c_int, ctypes.c_int]
    L.seqalign_shared_alloc.argtypes = [ctypes.c_int, sz, ctypes.POINTER(vp), vp]
    L.seqalign_shared_free.argtypes = [ctypes.c_int, vp]
    L.seqalign_shared_open.argtypes = [ctypes.c_int, vp, ctypes.POINTER(vp)]
    L.seqalign_shared_close.argtypes = [ctypes.c_int, vp]
    L.seqalign_last_create_error.restype = ctypes.c_char_p
    L.seqalign_batch_create.restype = vp
    L.seqalign_batch_create.argtypes = [ctypes.c_int]
    L.seqalign_batch_destroy.argtypes = [vp]
    L.seqalign_batch_error.restype = ctypes.c_char_p
    L.seqalign_batch_error.argtypes = [vp]
    L.seqalign_batch_set_scoring.argtypes = [vp, vp]
    L.seqalign_batch_submit.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, sz]
    L.seqalign_batch_submit_packed.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, sz]
    L.seqalign_batch_submit_uniform.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, sz, vp, sz, sz]
    L.seqalign_batch_set_result_sink.argtypes = [vp, vp, vp, vp]
    L.seqalign_batch_scores.argtypes = [vp, vp]
    L.seqalign_batch_ends.argtypes = [vp, vp, vp, vp]
    L.seqalign_batch_size.restype = sz
    L.seqalign_batch_size.argtypes = [vp]
    L.seqalign_batch_alignment.argtypes = [vp, sz, vp]
    L.seqalign_batch_run_device.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, sz, vp, vp, vp, vp]
    L.seqalign_batch_run_device_async.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, sz, vp, vp, vp, vp]
    L.seqalign_batch_run_device_wait.argtypes = [vp]
    L.seqalign_fill_matrices.argtypes = [vp, vp, sz, vp, sz, ctypes.c_int, vp, vp, vp]
    L.seqalign_batch_unknown_pair.argtypes = [vp, vp, vp]
    L.seqalign_batch_last_kernel_ms.restype = ctypes.c_double
    L.seqalign_batch_last_kernel_ms.argtypes = [vp]
    L.seqalign_batch_last_walk_ms.restype = ctypes.c_double
    L.seqalign_batch_last_walk_ms.argtypes = [vp]
    L.seqalign_batch_last_launches.restype = ctypes.c_int
    L.seqalign_batch_last_launches.argtypes = [vp]
    L.seqalign_batch_last_kernel.restype = ctypes.c_char_p
    L.seqalign_batch_last_kernel.argtypes = [vp]
    L.seqalign_batch_force_general.argtypes = [vp, ctypes.c_int]
    L.seqalign_batch_speculation_stats.argtypes = [vp, vp, vp]
    L.seqalign_batch_set_hit_limits.argtypes = [vp, sz, ctypes.c_int32]
    L.seqalign_batch_hit_count.restype = sz
    L.seqalign_batch_hit_count.argtypes = [vp, sz]
    L.seqalign_batch_hit.argtypes = [vp, sz, sz, vp]
    L.seqalign_batch_matrices.argtypes = [vp, sz, vp, vp, vp]
    L.seqalign_host_alloc.restype = vp
    L.seqalign_host_alloc.argtypes = [sz]
    L.seqalign_host_free.argtypes = [vp]
    L.seqalign_multi_create.restype = vp
    L.seqalign_multi_create.argtypes = [vp, ctypes.c_int]
    L.seqalign_multi_destroy.argtypes = [vp]
    L.seqalign_multi_devices.argtypes = [vp]
    L.seqalign_multi_error.restype = ctypes.c_char_p
    L.seqalign_multi_error.argtypes = [vp]
    L.seqalign_multi_set_scoring.argtypes = [vp, vp]
    L.seqalign_multi_set_hit_limits.argtypes = [vp, sz, ctypes.c_int32]
    L.seqalign_multi_submit_packed.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, sz]
    L.seqalign_multi_submit_uniform.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, sz, vp, sz, sz]
    L.seqalign_multi_size.restype = sz
    L.seqalign_multi_size.argtypes = [vp]
    L.seqalign_multi_scores.argtypes = [vp, vp]
    L.seqalign_multi_ends.argtypes = [vp, vp, vp, vp]
    L.seqalign_multi_alignment.argtypes = [vp, sz, vp]
    L.seqalign_multi_hit_count.restype = sz
    L.seqalign_multi_hit_count.argtypes = [vp, sz]
    L.seqalign_multi_hit.argtypes = [vp, sz, sz, vp]
    L.seqalign_multi_matrices.argtypes = [vp, sz, vp, vp, vp]
    L.seqalign_multi_where.argtypes = [vp, sz, vp, vp]
    L.seqalign_multi_last_kernel_ms.restype = ctypes.c_double
    L.seqalign_multi_last_kernel_ms.argtypes = [vp]
    L.seqalign_synth_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_int, ctypes.c_int, vp, vp, vp]
    # sequence-file text -> records on the device (csrc/sa_decode.cu)
    L.seqalign_reads_create.restype = vp
    L.seqalign_reads_create.argtypes = [ctypes.c_int]
    L.seqalign_reads_destroy.argtypes = [vp]
    L.seqalign_reads_error.restype = ctypes.c_char_p
    L.seqalign_reads_error.argtypes = [vp]
    L.seqalign_reads_decode.argtypes = [vp, vp, sz, ctypes.c_int, ctypes.c_int]
    L.seqalign_reads_format.argtypes = [vp]
    L.seqalign_reads_records.restype = sz
    L.seqalign_reads_records.argtypes = [vp]
    L.seqalign_reads_count.restype = sz
    L.seqalign_reads_count.argtypes = [vp, ctypes.c_int]
    L.seqalign_reads_offsets.restype = ctypes.POINTER(ctypes.c_int64)
    L.seqalign_reads_offsets.argtypes = [vp, ctypes.c_int]
    L.seqalign_reads_record_start.restype = sz
    L.seqalign_reads_record_start.argtypes = [vp, sz]
    L.seqalign_reads_name.argtypes = [vp, sz, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    L.seqalign_reads_fetch.argtypes = [vp, ctypes.c_int, vp]
    L.seqalign_reads_device_seq.restype = vp
    L.seqalign_reads_device_seq.argtypes = [vp, ctypes.c_int]
    L.seqalign_reads_device_offsets.restype = vp
    L.seqalign_reads_device_offsets.argtypes = [vp, ctypes.c_int]
    L.seqalign_reads_last_ms.restype = ctypes.c_double
    L.seqalign_reads_last_ms.argtypes = [vp]
    L.seqalign_batch_submit_reads.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, sz, sz]
    # reference C API
    L.scoring_init.argtypes = [vp] + [ctypes.c_int] * 4 + [ctypes.c_bool] * 6
    L.scoring_add_wildcard.argtypes = [vp, ctypes.c_char, ctypes.c_int]
    L.scoring_add_mutation.argtypes = [vp, ctypes.c_char, ctypes.c_char, ctypes.c_int]
    L.scoring_add_mutations.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_char]
    L.scoring_lookup.argtypes = [vp, ctypes.c_char, ctypes.c_char, vp, vp]
    for name in ("PAM30", "PAM70", "BLOSUM80", "BLOSUM62", "DNA_hybridization", "default"):
        getattr(L, "scoring_system_" + name).argtypes = [vp]
    L.alignment_create.restype = ctypes.POINTER(AlignmentT)
    L.alignment_create.argtypes = [sz]
    L.alignment_free.argtypes = [ctypes.POINTER(AlignmentT)]
    L.needleman_wunsch_new.restype = ctypes.POINTER(AlignerT)
    L.needleman_wunsch_free.argtypes = [ctypes.POINTER(AlignerT)]
    L.needleman_wunsch_align2.argtypes = [vp, vp, sz, sz, vp, ctypes.POINTER(AlignerT), ctypes.POINTER(AlignmentT)]
    L.smith_waterman_new.restype = vp
    L.smith_waterman_free.argtypes = [vp]
    L.smith_waterman_get_aligner.restype = ctypes.POINTER(AlignerT)
    L.smith_waterman_get_aligner.argtypes = [vp]
    L.smith_waterman_align2.argtypes = [vp, vp, sz, sz, vp, vp]
    L.smith_waterman_fetch.argtypes = [vp, ctypes.POINTER(AlignmentT)]
    L.aligner_align.argtypes = [ctypes.POINTER(AlignerT), vp, vp, sz, sz, vp, ctypes.c_char]
    L.aligner_destroy.argtypes = [ctypes.POINTER(AlignerT)]
    _lib = L
    return L


def device_count():
    return load().seqalign_device_count()


class Scoring:
    """scoring_t with the reference's constructors.

    ``Scoring(match, mismatch, gap_open, gap_extend, ...)`` is scoring_init;
    ``Scoring.system("BLOSUM62")`` the built-in systems; ``poke`` overwrites
    fields in place the way the CLI does (reference
    src/alignment_cmdline.c:401-439, src/tools/sw_cmdline.c:42-45), leaving
    min_penalty/max_penalty untouched.
    """

    def __init__(self, match=1, mismatch=-2, gap_open=-4, gap_extend=-1,
                 no_start_gap_penalty=False, no_end_gap_penalty=False,
                 no_gaps_in_a=False, no_gaps_in_b=False,
                 no_mismatches=False, case_sensitive=False):
        self.s = ScoringT()
        load().scoring_init(ctypes.byref(self.s), match, mismatch, gap_open, gap_extend,
                            no_start_gap_penalty, no_end_gap_penalty, no_gaps_in_a,
                            no_gaps_in_b, no_mismatches, case_sensitive)

    @classmethod
    def system(cls, name):
        self = cls.__new__(cls)
        self.s = ScoringT()
        getattr(load(), "scoring_system_" + name)(ctypes.byref(self.s))
        return self

    @classmethod
    def nw_default(cls):
        """needleman_wunsch default: 1/-2/-4/-1 (reference alignment_scoring.c:380-392)."""
        return cls.system("default")

    @classmethod
    def sw_cli_default(cls):
        """smith_waterman CLI default: default system poked to 2/-2/-2/-1
        (reference src/tools/sw_cmdline.c:37-46)."""
        return cls.system("default").poke(match=2, mismatch=-2, gap_open=-2, gap_extend=-1)

    def poke(self, **fields):
        for k, v in fields.items():
            setattr(self.s, k, v)
        return self

    def add_wildcard(self, c, score):
        load().scoring_add_wildcard(ctypes.byref(self.s), c.encode() if isinstance(c, str) else c, score)
        return self

    def add_mutation(self, a, b, score):
        enc = lambda c: c.encode() if isinstance(c, str) else c
        load().scoring_add_mutation(ctypes.byref(self.s), enc(a), enc(b), score)
        return self

    def add_mutations(self, letters, scores, use_match_mismatch=True):
        n = len(letters)
        arr = (ctypes.c_int * (n * n))(*[int(v) for v in scores])
        load().scoring_add_mutations(ctypes.byref(self.s), letters.encode(), arr,
                                     bytes([1 if use_match_mismatch else 0]))
        return self

    def lookup(self, a, b):
        score = ctypes.c_int()
        is_match = ctypes.c_bool()
        enc = lambda c: c.encode() if isinstance(c, str) else c
        load().scoring_lookup(ctypes.byref(self.s), enc(a), enc(b), ctypes.byref(score), ctypes.byref(is_match))
        return score.value, bool(is_match.value)

    @property
    def ptr(self):
        return ctypes.byref(self.s)


def pack(seqs):
    """list of bytes/str -> (uint8 array, int64 offsets[n+1])"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return data, off


class Alignment:
    __slots__ = ("result_a", "result_b", "score", "pos_a", "pos_b", "len_a", "len_b", "length")

    def __init__(self, al):
        n = al.length
        self.result_a = ctypes.string_at(al.result_a, n)
        self.result_b = ctypes.string_at(al.result_b, n)
        self.length = n
        self.score = al.score
        self.pos_a, self.pos_b, self.len_a, self.len_b = al.pos_a, al.pos_b, al.len_a, al.len_b

    def __repr__(self):
        return "Alignment(%r, %r, score=%d)" % (self.result_a, self.result_b, self.score)


class BatchAligner:
    """One engine on one device (seqalign_batch_t)."""

    def __init__(self, device=0, scoring=None):
        L = load()
        self._L = L
        self._h = L.seqalign_batch_create(device)
        if not self._h:
            raise SeqAlignError(ERR_CUDA, L.seqalign_last_create_error().decode())
        self._res = L.alignment_create(256)
        self._keep = None
        if scoring is not None:
            self.set_scoring(scoring)

    def close(self):
        if self._h:
            self._L.alignment_free(self._res)
            self._L.seqalign_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise SeqAlignError(rc, self._L.seqalign_batch_error(self._h).decode())
        return rc

    def set_scoring(self, scoring):
        self._check(self._L.seqalign_batch_set_scoring(self._h, scoring.ptr))

    def force_general(self, on=True):
        """0/False automatic, 1/True general kernel, 2 per-column end keys, 3 no end cell"""
        self._L.seqalign_batch_force_general(self._h, int(on))

    def submit_packed(self, algo, mode, seq_a, off_a, seq_b, off_b):
        """Host arrays (numpy uint8 / int64, or anything exposing ctypes.data)."""
        n = len(off_a) - 1
        self._keep = (seq_a, off_a, seq_b, off_b)
        self._check(self._L.seqalign_batch_submit_packed(
            self._h, algo, mode, seq_a.ctypes.data, off_a.ctypes.data,
            seq_b.ctypes.data, off_b.ctypes.data, n))
        return n

    def submit_ptrs(self, algo, mode, ptr_a, off_a_ptr, ptr_b, off_b_ptr, n):
        """Raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        self._check(self._L.seqalign_batch_submit_packed(self._h, algo, mode, ptr_a, off_a_ptr, ptr_b, off_b_ptr, n))
        return n

    def submit_reads(self, algo, mode, reads_a, side_a, reads_b, side_b, n, first=0):
        """align records [first, first + n) of two decoded sides where they lie in HBM (seqalign_batch_submit_reads)"""
        self._check(self._L.seqalign_batch_submit_reads(self._h, algo, mode, reads_a._h, side_a, reads_b._h, side_b, first, n))
        return n

    def submit_uniform_ptrs(self, algo, mode, ptr_a, len_a, ptr_b, len_b, n):
        """fixed-length pairs from raw host pointers: no offset arrays (seqalign_batch_submit_uniform)"""
        self._check(self._L.seqalign_batch_submit_uniform(self._h, algo, mode, ptr_a, len_a, ptr_b, len_b, n))
        return n

    def set_result_sink(self, score_ptr=0, x_ptr=0, y_ptr=0):
        """host pointers (ints) that receive the score-mode results of the following submits; 0 resets"""
        self._check(self._L.seqalign_batch_set_result_sink(self._h, score_ptr or None, x_ptr or None, y_ptr or None))

    def submit(self, algo, mode, seqs_a, seqs_b):
        a, oa = pack(seqs_a)
        b, ob = pack(seqs_b)
        return self.submit_packed(algo, mode, a, oa, b, ob)

    def scores(self):
        n = self._L.seqalign_batch_size(self._h)
        out = np.zeros(n, dtype=np.int32)
        self._check(self._L.seqalign_batch_scores(self._h, out.ctypes.data))
        return out

    def ends(self):
        n = self._L.seqalign_batch_size(self._h)
        s, x, y = (np.zeros(n, dtype=np.int32) for _ in range(3))
        self._check(self._L.seqalign_batch_ends(self._h, s.ctypes.data, x.ctypes.data, y.ctypes.data))
        return s, x, y

    def alignment(self, i):
        """Alignment of pair i (None for an SW pair without a hit)."""
        rc = self._check(self._L.seqalign_batch_alignment(self._h, i, self._res))
        return Alignment(self._res.contents) if rc == 1 else None

    def set_hit_limits(self, max_hits=8, min_score=1):
        self._check(self._L.seqalign_batch_set_hit_limits(self._h, max_hits, min_score))

    def hits(self, i):
        """all hits of pair i of the last MODE_HITS submit, in the reference's order"""
        out = []
        for h in range(self._L.seqalign_batch_hit_count(self._h, i)):
            self._check(self._L.seqalign_batch_hit(self._h, i, h, self._res))
            out.append(Alignment(self._res.contents))
        return out

    def run_device(self, algo, d_seq_a, d_off_a, d_seq_b, d_off_b, n, d_score, d_xend=0, d_yend=0, stream=0):
        """Device pointers (ints); results stay on the device."""
        self._check(self._L.seqalign_batch_run_device(self._h, algo, d_seq_a, d_off_a, d_seq_b, d_off_b,
                                                       n, d_score, d_xend or None, d_yend or None,
                                                       stream or None))

    def run_device_async(self, algo, d_seq_a, d_off_a, d_seq_b, d_off_b, n, d_score, d_xend=0, d_yend=0, stream=0):
        """run_device without waiting (up to 4 outstanding); pair every call with run_device_wait()"""
        self._check(self._L.seqalign_batch_run_device_async(self._h, algo, d_seq_a, d_off_a, d_seq_b, d_off_b,
                                                             n, d_score, d_xend or None, d_yend or None,
                                                             stream or None))

    def run_device_wait(self):
        """completes the oldest outstanding run_device_async"""
        self._check(self._L.seqalign_batch_run_device_wait(self._h))

    def fill_matrices(self, a, b, is_sw):
        a = a.encode() if isinstance(a, str) else bytes(a)
        b = b.encode() if isinstance(b, str) else bytes(b)
        cells = (len(a) + 1) * (len(b) + 1)
        m, ga, gb = (np.zeros(cells, dtype=np.int32) for _ in range(3))
        self._check(self._L.seqalign_fill_matrices(self._h, a, len(a), b, len(b), 1 if is_sw else 0,
                                                    m.ctypes.data, ga.ctypes.data, gb.ctypes.data))
        shape = (len(b) + 1, len(a) + 1)
        return m.reshape(shape), ga.reshape(shape), gb.reshape(shape)

    def matrices(self, i, len_a, len_b):
        """MODE_MATS: (match, gap_a, gap_b) of pair i as (len_b+1, len_a+1) int32 arrays"""
        cells = (len_a + 1) * (len_b + 1)
        m, ga, gb = (np.empty(cells, dtype=np.int32) for _ in range(3))
        self._check(self._L.seqalign_batch_matrices(self._h, i, m.ctypes.data, ga.ctypes.data, gb.ctypes.data))
        shape = (len_b + 1, len_a + 1)
        return m.reshape(shape), ga.reshape(shape), gb.reshape(shape)

    def speculation_stats(self):
        h, m = ctypes.c_int(), ctypes.c_int()
        self._L.seqalign_batch_speculation_stats(self._h, ctypes.byref(h), ctypes.byref(m))
        return h.value, m.value

    @property
    def last_kernel_ms(self):
        return self._L.seqalign_batch_last_kernel_ms(self._h)

    @property
    def last_walk_ms(self):
        return self._L.seqalign_batch_last_walk_ms(self._h)

    @property
    def last_launches(self):
        return self._L.seqalign_batch_last_launches(self._h)

    @property
    def last_kernel(self):
        return self._L.seqalign_batch_last_kernel(self._h).decode()


class MultiAligner:
    """One batch over several devices of this node from one process (seqalign_multi_t): pair ranges
    balanced by cells, one engine and host thread per device, results by global pair index."""

    def __init__(self, devices=None, scoring=None):
        L = load()
        self._L = L
        if devices is None:
            self._h = L.seqalign_multi_create(None, 0)
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            self._h = L.seqalign_multi_create(arr, len(devices))
        if not self._h:
            raise SeqAlignError(ERR_CUDA, L.seqalign_last_create_error().decode() or "no usable device")
        self._res = L.alignment_create(256)
        self._keep = None
        if scoring is not None:
            self.set_scoring(scoring)

    def close(self):
        if self._h:
            self._L.alignment_free(self._res)
            self._L.seqalign_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise SeqAlignError(rc, self._L.seqalign_multi_error(self._h).decode())
        return rc

    @property
    def devices(self):
        return self._L.seqalign_multi_devices(self._h)

    def set_scoring(self, scoring):
        self._check(self._L.seqalign_multi_set_scoring(self._h, scoring.ptr))

    def set_hit_limits(self, max_hits=8, min_score=1):
        self._check(self._L.seqalign_multi_set_hit_limits(self._h, max_hits, min_score))

    def submit_packed(self, algo, mode, seq_a, off_a, seq_b, off_b):
        self._keep = (seq_a, off_a, seq_b, off_b)
        n = len(off_a) - 1
        self._check(self._L.seqalign_multi_submit_packed(self._h, algo, mode, seq_a.ctypes.data, off_a.ctypes.data,
                                                         seq_b.ctypes.data, off_b.ctypes.data, n))
        return n

    def submit_uniform(self, algo, mode, seq_a, len_a, seq_b, len_b, n):
        self._keep = (seq_a, seq_b)
        self._check(self._L.seqalign_multi_submit_uniform(self._h, algo, mode, seq_a.ctypes.data, len_a, seq_b.ctypes.data, len_b, n))
        return n

    def ends(self):
        n = self._L.seqalign_multi_size(self._h)
        s, x, y = (np.zeros(n, dtype=np.int32) for _ in range(3))
        self._check(self._L.seqalign_multi_ends(self._h, s.ctypes.data, x.ctypes.data, y.ctypes.data))
        return s, x, y

    def scores(self):
        n = self._L.seqalign_multi_size(self._h)
        s = np.zeros(n, dtype=np.int32)
        self._check(self._L.seqalign_multi_scores(self._h, s.ctypes.data))
        return s

    def alignment(self, i):
        rc = self._check(self._L.seqalign_multi_alignment(self._h, i, self._res))
        return Alignment(self._res.contents) if rc == 1 else None

    def hits(self, i):
        out = []
        for h in range(self._L.seqalign_multi_hit_count(self._h, i)):
            self._check(self._L.seqalign_multi_hit(self._h, i, h, self._res))
            out.append(Alignment(self._res.contents))
        return out

    def where(self, i):
        d, l = ctypes.c_int(), ctypes.c_size_t()
        self._check(self._L.seqalign_multi_where(self._h, i, ctypes.byref(d), ctypes.byref(l)))
        return d.value, l.value


def needleman_wunsch(a, b, scoring):
    """needleman_wunsch_align through the C API (single pair)."""
    L = load()
    a = a.encode() if isinstance(a, str) else bytes(a)
    b = b.encode() if isinstance(b, str) else bytes(b)
    nw = L.needleman_wunsch_new()
    res = L.alignment_create(256)
    try:
        L.needleman_wunsch_align2(a, b, len(a), len(b), scoring.ptr, nw, res)
        return Alignment(res.contents)
    finally:
        L.alignment_free(res)
        L.needleman_wunsch_free(nw)


def smith_waterman(a, b, scoring, max_hits=None):
    """smith_waterman_align + fetch loop through the C API (single pair)."""
    L = load()
    a = a.encode() if isinstance(a, str) else bytes(a)
    b = b.encode() if isinstance(b, str) else bytes(b)
    sw = L.smith_waterman_new()
    res = L.alignment_create(256)
    hits = []
    try:
        L.smith_waterman_align2(a, b, len(a), len(b), scoring.ptr, sw)
        while (max_hits is None or len(hits) < max_hits) and L.smith_waterman_fetch(sw, res):
            hits.append(Alignment(res.contents))
        return hits
    finally:
        L.alignment_free(res)
        L.smith_waterman_free(sw)


class PipelinedAligner:
    """Keeps `depth` batches in flight on one device: each worker thread owns an
    engine (the C-ABI is one engine per host thread), so the PCIe copy of one
    batch overlaps the DP kernel of another.  ctypes releases the GIL during
    the calls.  `map_scores` yields score arrays in submission order."""

    def __init__(self, device=0, scoring=None, depth=2):
        from concurrent.futures import ThreadPoolExecutor
        import queue
        self._engines = queue.Queue()
        self._all = [BatchAligner(device, scoring) for _ in range(depth)]
        for e in self._all:
            self._engines.put(e)
        self._pool = ThreadPoolExecutor(max_workers=depth)
        self.depth = depth

    def _run(self, algo, mode, ptrs, n):
        eng = self._engines.get()
        try:
            eng.submit_ptrs(algo, mode, *ptrs, n)
            return eng.scores() if mode == MODE_SCORE_ONLY else eng.ends()
        finally:
            self._engines.put(eng)

    def submit_ptrs(self, algo, mode, ptr_a, off_a_ptr, ptr_b, off_b_ptr, n):
        """asynchronous: returns a future whose result() is scores (MODE_SCORE_ONLY)
        or (score, x_end, y_end)"""
        return self._pool.submit(self._run, algo, mode, (ptr_a, off_a_ptr, ptr_b, off_b_ptr), n)

    def _run_uniform(self, algo, mode, ptr_a, len_a, ptr_b, len_b, n, sink):
        eng = self._engines.get()
        try:
            eng.set_result_sink(sink)
            eng.submit_uniform_ptrs(algo, mode, ptr_a, len_a, ptr_b, len_b, n)
            return eng.last_kernel_ms
        finally:
            eng.set_result_sink(0)
            self._engines.put(eng)

    def submit_uniform_ptrs(self, algo, mode, ptr_a, len_a, ptr_b, len_b, n, score_sink):
        """fixed-length pairs from raw (pinned) host pointers, scores written straight to the host
        array at `score_sink` (n int32) by the device->host copy; the future's result is the kernel time"""
        return self._pool.submit(self._run_uniform, algo, mode, ptr_a, len_a, ptr_b, len_b, n, score_sink)

    def submit_packed(self, algo, mode, seq_a, off_a, seq_b, off_b):
        keep = (seq_a, off_a, seq_b, off_b)
        fut = self.submit_ptrs(algo, mode, seq_a.ctypes.data, off_a.ctypes.data, seq_b.ctypes.data,
                               off_b.ctypes.data, len(off_a) - 1)
        fut._keep = keep
        return fut

    def map_scores(self, algo, batches, mode=MODE_SCORE_ONLY):
        """batches: iterable of (seq_a, off_a, seq_b, off_b) numpy arrays"""
        import collections
        pending = collections.deque()
        for b in batches:
            pending.append(self.submit_packed(algo, mode, *b))
            if len(pending) > self.depth:
                yield pending.popleft().result()
        while pending:
            yield pending.popleft().result()

    @property
    def engines(self):
        return list(self._all)

    def close(self):
        self._pool.shutdown(wait=True)
        for e in self._all:
            e.close()


ERR_IRREGULAR = -6
FMT_PLAIN, FMT_FASTA, FMT_FASTQ = 1, 2, 4


class Reads:
    """Records of a chunk of sequence-file text, decoded on the device (seqalign_reads_*, csrc/sa_decode.cu):
    FASTA / FASTQ / one sequence per line as the reference's reader takes them (libs/seq_file/seq_file.h:245-325).
    decode() returns False when the text leaves the grammar the device takes (the host reader handles those)."""

    def __init__(self, device=0):
        self._L = load()
        self._h = self._L.seqalign_reads_create(device)
        if not self._h:
            raise SeqAlignError(-1, "cannot create a reads object on device %d (sm_100 only, no CPU path)" % device)
        self._text = b""

    def close(self):
        if self._h:
            self._L.seqalign_reads_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, text, final=True, split=False):
        self._text = bytes(text)
        buf = ctypes.create_string_buffer(self._text, len(self._text) + 1)
        rc = self._L.seqalign_reads_decode(self._h, buf, len(self._text), int(bool(final)), int(bool(split)))
        if rc == ERR_IRREGULAR:
            return False
        if rc < 0:
            raise SeqAlignError(rc, self._L.seqalign_reads_error(self._h).decode())
        return True

    @property
    def format(self):
        return self._L.seqalign_reads_format(self._h)

    @property
    def records(self):
        return self._L.seqalign_reads_records(self._h)

    @property
    def last_ms(self):
        return self._L.seqalign_reads_last_ms(self._h)

    def count(self, side=0):
        return self._L.seqalign_reads_count(self._h, side)

    def offsets(self, side=0):
        n = self.count(side)
        p = self._L.seqalign_reads_offsets(self._h, side)
        return np.array([p[i] for i in range(n + 1)], dtype=np.int64) if p else np.zeros(1, dtype=np.int64)

    def record_start(self, i):
        return self._L.seqalign_reads_record_start(self._h, i)

    def name(self, i):
        pos, ln = ctypes.c_size_t(0), ctypes.c_size_t(0)
        rc = self._L.seqalign_reads_name(self._h, i, ctypes.byref(pos), ctypes.byref(ln))
        if rc < 0:
            raise SeqAlignError(rc, "bad record index")
        return self._text[pos.value:pos.value + ln.value]

    def sequences(self, side=0):
        """the side's sequences as a list of bytes (complete records only)"""
        off = self.offsets(side)
        total = int(off[-1]) if len(off) else 0
        # the packed buffer also holds the held-back tail record: fetch copies all of it
        out = ctypes.create_string_buffer(len(self._text) + 64)
        rc = self._L.seqalign_reads_fetch(self._h, side, out)
        if rc < 0:
            raise SeqAlignError(rc, self._L.seqalign_reads_error(self._h).decode())
        raw = out.raw[:total]
        return [raw[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def synth_device(device, kind, seed, first_pair, npairs, len_a, len_b, d_seq_a, d_seq_b, stream=0):
    """pairs [first_pair, first_pair+npairs) of the synthetic stream `seed` written into device
    buffers (pointers as ints) by the CUDA generator; same bytes as seqalign.synth.synth_batch"""
    from .synth import KINDS
    rc = load().seqalign_synth_batch(int(device), KINDS[kind][0], int(seed), int(first_pair), int(npairs),
                                     int(len_a), int(len_b), d_seq_a, d_seq_b, stream or None)
    if rc != 0:
        raise SeqAlignError(rc, "seqalign_synth_batch failed")


def enable_peer_access(device, peer):
    """kernels of `device` may read memory of `peer` (NVLink); see seqalign_enable_peer_access"""
    rc = load().seqalign_enable_peer_access(int(device), int(peer))
    if rc != 0:
        raise SeqAlignError(rc, "device %d cannot access device %d as a peer" % (device, peer))


class SharedBuffer:
    """Device memory one process owns and its peers (one process per GPU, same
    node) map onto their own device (seqalign_shared_alloc / _open).  `handle`
    is 64 opaque bytes to hand to the peers; `tensor()` views the memory as a
    torch uint8 tensor without copying."""

    def __init__(self, device, nbytes=None, handle=None):
        L = load()
        self.device, self.ptr = int(device), ctypes.c_void_p()
        self.owner = handle is None
        if self.owner:
            buf = ctypes.create_string_buffer(64)
            rc = L.seqalign_shared_alloc(self.device, int(nbytes), ctypes.byref(self.ptr), buf)
            self.handle, self.nbytes = buf.raw, int(nbytes)
        else:
            rc = L.seqalign_shared_open(self.device, ctypes.create_string_buffer(handle, 64), ctypes.byref(self.ptr))
            self.handle, self.nbytes = handle, int(nbytes)
        if rc != 0:
            raise SeqAlignError(rc, "cannot %s a shared device buffer on device %d" % ("allocate" if self.owner else "open", self.device))

    @property
    def address(self):
        return self.ptr.value

    def tensor(self):
        import torch

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = dict(shape=(self.nbytes,), typestr="|u1", data=(self.ptr.value, False), version=2)
        return torch.as_tensor(v, device="cuda:%d" % self.device)

    def close(self):
        if self.ptr.value:
            L = load()
            (L.seqalign_shared_free if self.owner else L.seqalign_shared_close)(self.device, self.ptr)
            self.ptr = ctypes.c_void_p()

