/*
 * mmannot_b200 -- C ABI of the B200 read-annotation hot path.
 *
 * This is the drop-in boundary for the part of mzytnicki/mmannot that annotates
 * alignment records ("hits") and counts element combinations.  The reference has
 * no FFI of its own; the entry points below sit exactly where its hot loop calls
 *
 *     IntervalList::scan(Read&, regions, ids, position)      mmannot.cpp:1291-1332
 *     Counter::addCount(name, regions, ids, nHits)           mmannot.cpp:1665-1739
 *     Counter::read() end-of-file flush                      mmannot.cpp:1783-1800
 *     Counter::getCounts() / TableCount::addCounter()        mmannot.cpp:1803, 1861-1876
 *
 * (called from Counter::read, mmannot.cpp:1772-1778, and main, mmannot.cpp:2109-2115).
 * Everything before that loop (config, GTF -> typed intervals, SAM/BAM decode) and
 * after it (table / statistics formatting) stays on the host.
 *
 * Conventions: plain pointers and sizes, no C++ types, no exceptions, no exit().
 * Every function returns MMA_OK (0) or a negative MMA_ERR_* code; the message is
 * available from mma_last_error().  Calls on one context must be serialised by the
 * caller; different contexts (one per GPU) may be driven from different threads.
 * There is no CPU fallback: without a usable CUDA device mma_create() fails.
 */
#ifndef MMANNOT_B200_H
#define MMANNOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMA_OK 0
#define MMA_ERR_INVALID (-1)   /* bad argument / unsorted features / unsupported size */
#define MMA_ERR_CUDA (-2)      /* CUDA runtime error */
#define MMA_ERR_STATE (-3)     /* wrong call sequence */
#define MMA_ERR_CAPACITY (-4)  /* combination table or staging capacity exceeded */
#define MMA_ERR_NO_DEVICE (-5) /* no CUDA device: the product path has no CPU fallback */
#define MMA_ERR_RETRY (-6)     /* an optimistic multi-GPU exchange met a shard with deferred records: see mma_export_table_async */

/* -y (Strategy, mmannot.cpp:50, 2033-2045) */
#define MMA_STRATEGY_DEFAULT 0
#define MMA_STRATEGY_UNIQUE 1
#define MMA_STRATEGY_RANDOM 2
#define MMA_STRATEGY_RATIO 3

/* element strand / feature strand (Strand, mmannot.cpp:49) */
#define MMA_STRAND_ALL 0
#define MMA_STRAND_F 1
#define MMA_STRAND_R 2

/* element kind used for the distance tie-break (Config::isUpstream/isDownstream, mmannot.cpp:463-470) */
#define MMA_VICINITY_NONE 0
#define MMA_VICINITY_UP 1
#define MMA_VICINITY_DOWN 2

/* hit meta word */
#define MMA_HIT_CHR_MASK 0x00FFFFFFu   /* bits 0..23: annotation chromosome id */
#define MMA_HIT_CHR_NONE 0x00FFFFFFu   /* chromosome unknown to the annotation (mmannot.cpp:1294-1302) */
#define MMA_HIT_STRAND_BIT 0x80000000u /* bit 31: read strand after the -s mapping (mmannot.cpp:836-844, 884) */

#define MMA_MAX_ELEMENTS 64
#define MMA_FAST_OFF 0xFFFFFFFFu

typedef struct mma_ctx mma_ctx; /* one per GPU, opaque */

typedef struct mma_params {
  int32_t device;          /* CUDA device ordinal */
  int32_t strategy;        /* MMA_STRATEGY_* */
  float overlap;           /* -l as parsed by stof (Globals::overlap, mmannot.cpp:70, 1972-1977):
                              <0 inclusion, <1 fraction of the read, else nucleotides */
  float rescue_threshold;  /* -e/100 in fp32 (mmannot.cpp:2024); >=1 disables rescue */
  int32_t read_stats;      /* -m: rescue only acts when read statistics are requested (mmannot.cpp:491) */
  int32_t interval_stats;  /* -M */
  uint32_t n_elements;     /* E = flattened length of the Order section, <= MMA_MAX_ELEMENTS */
  const uint16_t *elem_line;    /* [E] Order line (priority rank) of each element */
  const uint8_t *elem_strand;   /* [E] MMA_STRAND_* of the element (' +' / ' -' suffix) */
  const uint8_t *elem_vicinity; /* [E] MMA_VICINITY_* */
  uint32_t n_samples;      /* number of input files (table columns) */
  uint32_t max_batch_hits; /* staging capacity: largest n of one mma_submit_hits call */
  uint32_t table_log2;     /* log2(slots) of the per-sample combination table; 0 = 16 */
  uint32_t bin_shift;      /* log2(bin width) of the position index; 0 = chosen from the annotation extent */
  uint32_t rand_seed;      /* -y random: seed of the glibc rand() stream the reference draws from (1 = unseeded) */
  uint32_t fast_bin_shift; /* log2(bin width) of the segment answer table; 0 = chosen from the annotation, MMA_FAST_OFF = no table */
} mma_params;

/* Typed intervals in REFERENCE ORDER (mmannot.cpp:1267: sorted by chromosome id then start,
 * ties in the order the reference's std::sort leaves them).  1-based closed coordinates. */
typedef struct mma_features {
  uint32_t n;
  uint32_t n_chr;          /* number of annotation chromosomes; chr[i] < n_chr */
  const uint32_t *chr;
  const uint32_t *start;
  const uint32_t *end;
  const uint8_t *type;     /* flattened Order element index */
  const uint8_t *strand;   /* MMA_STRAND_F or MMA_STRAND_R */
} mma_features;

/* One batch of hits, struct-of-arrays, in file order.  start/end are the reference's
 * Read interval (mmannot.cpp:852-886: end = start + sum(M,D,=,X) - 1). */
typedef struct mma_hit_batch {
  uint64_t n;
  const uint32_t *start;
  const uint32_t *end;
  const uint32_t *meta;      /* MMA_HIT_* */
  const uint32_t *nh;        /* NH of the record (XamRecord::nHits) */
  const uint64_t *read_key;  /* 64-bit key of the read name (the reference keys by the name string) */
} mma_hit_batch;

/* The same batch in the compact transfer format (8 bytes per hit + 8 bytes per run of records sharing a read name, instead
 * of 24 bytes per hit): what crosses PCIe when the host decoder packs its output with mma_pack_hits().  The device expands
 * it back into the five arrays of mma_hit_batch (k_expand_packed) before the batch kernel runs; results are identical.
 *   packed[i]  bits 0..7   end - start + 1  (0 for the empty-CIGAR case end = start - 1; 255 = look the hit up in the escapes)
 *              bits 8..15  NH               (255 = look the hit up in the escapes)
 *              bits 16..29 annotation chromosome id (MMA_PACKED_CHR_NONE = unknown to the annotation)
 *              bit  30     first record of a run: its read key differs from the previous record's (always set for record 0)
 *              bit  31     read strand (MMA_HIT_STRAND_BIT)
 *   run_key[r]             read key of the r-th run of the batch
 *   tile_run_base[t]       number of runs that start before hit t * MMA_PACK_TILE
 *   esc_*                  ascending hit indices with their full end and NH, for hits a field of which did not fit
 * Normally the output of mma_pack_hits.  mma_submit_hits_packed checks what costs O(tiles + escapes) (mma_check_packed, deep = 0:
 * tile_run_base monotone and within n_runs, esc_index strictly increasing and below n) and the device never reads outside the
 * arrays; run-start bits that disagree with tile_run_base / n_runs, or an escaped hit missing from esc_index, give wrong read
 * keys / saturated fields without an error -- callers that build the struct themselves can run mma_check_packed(.., 1) on it. */
#define MMA_PACK_TILE 1024
#define MMA_PACKED_CHR_NONE 0x3FFFu
#define MMA_PACKED_RUN_START 0x40000000u
typedef struct mma_packed_batch {
  uint64_t n;
  const uint32_t *start;
  const uint32_t *packed;
  uint64_t n_runs;
  const uint64_t *run_key;
  const uint32_t *tile_run_base;  /* [(n + MMA_PACK_TILE - 1) / MMA_PACK_TILE] */
  uint64_t n_escapes;
  const uint32_t *esc_index;
  const uint32_t *esc_end;
  const uint32_t *esc_nh;
} mma_packed_batch;

typedef struct mma_sample_stats { /* Counter's counters, mmannot.cpp:1663, printed at 1807-1818 */
  uint64_t n_hits, n_reads, n_unique, n_ambiguous, n_multiple, n_unassigned, n_rescued;
} mma_sample_stats;

/* One row per distinct element combination.  The value the reference keeps in
 * regionCounts (a double, mmannot.cpp:1658) is  sum over rows with the same mask of
 * count * (nh ? 1.0 / nh : 1.0);  nh is non-zero only under -y ratio. */
typedef struct mma_sample_result {
  mma_sample_stats stats;
  uint64_t n_rows;
  const uint64_t *row_mask;  /* bit i set <=> element i in the combination */
  const uint32_t *row_nh;
  const uint64_t *row_count;
} mma_sample_result;

/* Per-kernel device time since mma_timing_reset(), from CUDA events recorded on the
 * context's own compute stream (timing must have been switched on). */
typedef struct mma_timing {
  double ms_index;    /* K1 feature index + segment answer table build */
  double ms_batch;    /* k_batch: K2 per-hit annotation + K3 per-read resolution + K4 counting, one kernel per batch */
  double ms_close;    /* k_batch_close: end-of-batch bookkeeping (+ re-routing of reads left unfinished mid-batch) */
  double ms_finish;   /* deferred (name-sorted) resolution at mma_finish_sample */
  uint64_t launches;  /* kernels launched by this library since the reset */
  uint64_t hits;      /* hits submitted since the reset */
  uint64_t batches;   /* k_batch launches since the reset */
  uint64_t fast_miss; /* hits the segment table could not answer (since the sample was reset) */
  double ms_bam_inflate; /* mma_submit_bam: BGZF inflate ... */
  double ms_bam_index;   /* ... record walk + prefix sum ... */
  double ms_bam_parse;   /* ... record parse into hits */
} mma_timing;

int mma_device_count(void); /* number of CUDA devices visible to the process (0 when there is none) */
int mma_warmup(int device); /* creates the CUDA context of `device` (a few 100 ms); callable from any thread, e.g. while the
                               annotation is being parsed */
int mma_create(mma_ctx **out, const mma_params *params);
void mma_destroy(mma_ctx *ctx);
const char *mma_last_error(const mma_ctx *ctx); /* ctx may be NULL: error of the last failed mma_create */

/* Uploads the feature buffer and builds the device index (chromosome offsets, running
 * max-end, position bins).  Replaces the sort/bins of mmannot.cpp:1267-1284. */
int mma_load_features(mma_ctx *ctx, const mma_features *features);

/* Page-locked host memory for hit buffers (plain malloc'ed buffers also work, slower). */
void *mma_alloc_pinned(size_t bytes);
void mma_free_pinned(void *p);

/* Asynchronous: copies the batch to the device on a side stream and enqueues the kernels.
 * The caller's buffers must stay untouched until the NEXT mma_submit_hits /
 * mma_finish_sample / mma_sync on this context returns (double-buffer on the host).
 * Replaces scan() + addCount() for the hits of the batch (mmannot.cpp:1772-1778). */
int mma_submit_hits(mma_ctx *ctx, uint32_t sample, const mma_hit_batch *batch);

/* Same, but the arrays already live in device memory of ctx's GPU (no copy is made; they
 * must stay valid until mma_sync / mma_finish_sample). */
int mma_submit_hits_device(mma_ctx *ctx, uint32_t sample, const mma_hit_batch *device_batch);

/* Host-side packer (no device work): fills the caller's buffers -- packed[n], run_key[n] (worst case one run per hit),
 * tile_run_base[(n + MMA_PACK_TILE - 1) / MMA_PACK_TILE], esc_*[esc_capacity] -- from a batch in the wide format and points
 * `out` at them (out->start aliases wide->start).  MMA_ERR_CAPACITY when the batch cannot be packed (more escapes than
 * esc_capacity, or a chromosome id >= MMA_PACKED_CHR_NONE): submit it in the wide format instead. */
int mma_pack_hits(const mma_hit_batch *wide, uint32_t *packed, uint64_t *run_key, uint32_t *tile_run_base, uint32_t *esc_index,
                  uint32_t *esc_end, uint32_t *esc_nh, uint64_t esc_capacity, mma_packed_batch *out);

/* Host-side consistency check of a packed batch (no device work, no context): MMA_OK or MMA_ERR_INVALID.  deep = 0: pointers,
 * counts, tile_run_base and esc_index, O(tiles + escapes); deep != 0: also the run-start bits of every tile and the escaped hits, O(n). */
int mma_check_packed(const mma_packed_batch *batch, int deep);

/* mma_submit_hits for a batch in the compact format (same asynchrony and buffer lifetime rules). */
int mma_submit_hits_packed(mma_ctx *ctx, uint32_t sample, const mma_packed_batch *batch);

/* ---- BAM decode on the device (BamReader, mmannot.cpp:1487-1649): the compressed BGZF members cross PCIe as they lie in the
 * file and are inflated and parsed into hits on the GPU, which then annotates them like any other batch.
 *   mma_bam_begin   once per input file, after the caller has read the BAM header (on the host: it is a few KB): the
 *                   annotation chromosome of every BAM reference (MMA_HIT_CHR_NONE = not in the annotation) and -s.
 *   mma_submit_bam  a chunk of WHOLE members (host memory, page-locked for speed) with their byte offsets and inflated sizes
 *                   (the ISIZE field of each member); skip_first = bytes at the start of the chunk's first member that are
 *                   not alignment records (the end of the BAM header; 0 for later chunks).  Returns with *n_records = records
 *                   (= hits) of the chunk and *flags = 0 once the chunk's kernels are enqueued.  A non-zero *flags (MMA_BAM_*)
 *                   means NOTHING of the chunk was counted: the file holds something this route leaves to the host decoder
 *                   (XA alternative hits, CIGAR operations or aux types the reference warns about, records that straddle
 *                   members, corrupt data, two neighbouring records whose names differ but hash to one read key); the caller
 *                   resets the sample and decodes the file itself (mma_submit_hits*).
 *   mma_bam_ref_first  out[i] = ordinal (0-based, over the file) of the first record on BAM reference i, ~0 if none yet --
 *                   kept only for references mapped to MMA_HIT_CHR_NONE (others stay ~0): for the "chromosome not present in
 *                   your annotation" warnings (mmannot.cpp:1297), in order of appearance. */
#define MMA_BAM_BAD_DEFLATE 1u
#define MMA_BAM_STRADDLE 2u
#define MMA_BAM_HAS_XA 4u
#define MMA_BAM_ODD_CIGAR 8u
#define MMA_BAM_ODD_AUX 16u
#define MMA_BAM_MALFORMED 32u
#define MMA_BAM_KEY_COLLISION 64u /* neighbouring records with one 64-bit read key and different names (mmannot.cpp:1656-1662 keys by the name) */
typedef struct mma_bam_chunk {
  const void *data;              /* whole BGZF members, back to back */
  uint64_t n_bytes;              /* < 2^32 */
  const uint32_t *member_offset; /* [n_members + 1] */
  const uint32_t *member_isize;  /* [n_members]; their sum < 2^32 */
  uint32_t n_members;
  uint32_t skip_first;
} mma_bam_chunk;
int mma_bam_begin(mma_ctx *ctx, uint32_t sample, const uint32_t *ref_to_chr, uint32_t n_ref, int strandedness /* 0 U, 1 F, 2 R */);
int mma_submit_bam(mma_ctx *ctx, uint32_t sample, const mma_bam_chunk *chunk, uint64_t *n_records, uint32_t *flags);
/* The same in two halves, so that the caller can read and stage the NEXT chunk while this one is inflated: _start enqueues the
 * inflate and the record walk and returns at once (the chunk's arrays may be released when it returns); _finish waits for them
 * and runs the record parse and the batch kernels, with the outputs of mma_submit_bam.  One chunk in flight at a time;
 * mma_bam_stage calls made between the two build the next chunk. */
int mma_submit_bam_start(mma_ctx *ctx, uint32_t sample, const mma_bam_chunk *chunk);
int mma_submit_bam_finish(mma_ctx *ctx, uint64_t *n_records, uint32_t *flags);
/* Staged upload, so that a large chunk (the inflate kernel wants tens of thousands of members per launch) can be fed from small
 * page-locked buffers while the file is still being read: mma_bam_stage copies n_bytes to byte `offset` of the chunk under
 * construction (asynchronously; the host buffer is free again when the NEXT mma_bam_stage / mma_submit_bam call returns, so two
 * buffers suffice), and mma_submit_bam with chunk->data == NULL takes the staged bytes [0, n_bytes) as its data.
 * mma_bam_reserve sizes the staging area up front (optional). */
int mma_bam_reserve(mma_ctx *ctx, uint64_t n_bytes);
int mma_bam_stage(mma_ctx *ctx, const void *host_bytes, uint64_t n_bytes, uint64_t offset);
int mma_bam_ref_first(mma_ctx *ctx, uint64_t *out, uint32_t n_ref);
/* The hits mma_submit_bam decoded from the last chunk, copied to host arrays of n_records entries (tests: the decoder against
 * the host's XamReader).  Synchronous. */
int mma_bam_last_hits(mma_ctx *ctx, uint32_t *start, uint32_t *end, uint32_t *meta, uint32_t *nh, uint64_t *read_key);

/* Synchronous: end-of-file flush of the sample (mmannot.cpp:1783-1792) and read-back of its
 * counters and rows.  The arrays belong to the context and stay valid until the next
 * mma_finish_sample / mma_reset_sample / mma_destroy. */
int mma_finish_sample(mma_ctx *ctx, uint32_t sample, mma_sample_result *out);

/* IntervalList::scan alone (mmannot.cpp:1291-1332): the element set (bit i <=> element i) of every hit of a HOST
 * batch, written to out_masks[0..n) (host memory).  Synchronous; nothing is counted; nh / read_key are ignored. */
int mma_annotate_hits(mma_ctx *ctx, const mma_hit_batch *batch, uint64_t *out_masks);

/* scan with -M (EvaluationStructure::getIds, mmannot.cpp:1077-1081, 1328-1330): besides the element set of every hit, the
 * indices (into the feature buffer given to mma_load_features) of the intervals behind it -- every interval of a chosen
 * element that passes the strand rule and the -l test.  out_offsets has n + 1 entries: the intervals of hit i are
 * (*out_ids)[out_offsets[i] .. out_offsets[i+1]), in no particular order (every consumer in the reference sorts them).
 * *out_ids belongs to the context and stays valid until the next call.  Synchronous; nothing is counted. */
int mma_annotate_intervals(mma_ctx *ctx, const mma_hit_batch *batch, uint64_t *out_masks, uint64_t *out_offsets, const uint32_t **out_ids);

/* Forget everything counted for `sample` (Counter::clear, mmannot.cpp:1742-1747). */
int mma_reset_sample(mma_ctx *ctx, uint32_t sample);

/* Dense device vector for a cross-GPU sum: out_dev[i] = count of row (mask[i], nh[i]) of
 * `sample`, 0 if absent.  out_dev is device memory (n x uint64); the collective itself is
 * issued by whoever owns the communicator (torch.distributed / NCCL). */
int mma_dense_counts(mma_ctx *ctx, uint32_t sample, const uint64_t *mask, const uint32_t *nh, uint64_t n, uint64_t *out_dev);

/* Multi-GPU merge of a sample on the devices (the sum TableCount::addCounter forms column by column, mmannot.cpp:1861-1876):
 *   1. every rank:  mma_export_table(ctx, sample, dev_buf)   end-of-file flush, then the compacted table and the counters
 *                   into dev_buf (device memory of ctx's GPU, mma_export_bytes(ctx) bytes), enqueued on mma_stream(ctx)
 *   2. the caller all-gathers the buffers (NCCL over NVLink; the communicator belongs to the caller)
 *   3. every rank:  mma_import_tables(ctx, sample, gathered, n_ranks)   replaces the sample's table and counters by the
 *                   sum over the n_ranks buffers (mma_export_bytes apart), on mma_stream(ctx)
 *   4. mma_finish_sample as usual: the merged result, identical on every rank.
 * Only the control block crosses PCIe before step 4. */
uint64_t mma_export_bytes(const mma_ctx *ctx);
int mma_export_table(mma_ctx *ctx, uint32_t sample, void *dev_dst);
int mma_import_tables(mma_ctx *ctx, uint32_t sample, const void *dev_src, uint32_t n_tables);
/* Rows of the last mma_export_table of this context, and the import of dumps exchanged at that size only: the dumps lie
 * stride_bytes apart (a multiple of 16, at least mma_export_head_bytes() + 16 * rows_cap) and hold at most rows_cap rows each,
 * so that the all-gather moves the live rows (a few thousand) instead of the whole table capacity. */
uint64_t mma_export_rows(const mma_ctx *ctx);
uint64_t mma_export_head_bytes(void);
int mma_import_tables_strided(mma_ctx *ctx, uint32_t sample, const void *dev_src, uint32_t n_tables, uint64_t stride_bytes, uint64_t rows_cap);

/* mma_export_table without any host synchronisation: flush, compaction and the copy of the head and the first rows_cap rows
 * (stride_bytes in all, see mma_import_tables_strided) are only ENQUEUED on mma_stream(ctx), so that the exchange and the import
 * can be queued behind the batch kernels while they still run.  The price: a shard that holds deferred records (input whose
 * reads are not adjacent) cannot resolve them first; the import then marks the merged sample, mma_finish_sample returns
 * MMA_ERR_RETRY on every rank, and each rank calls mma_restore_export (its own table and counters back from the dump it still
 * holds) and repeats the exchange with mma_export_table. */
int mma_export_table_async(mma_ctx *ctx, uint32_t sample, void *dev_dst, uint64_t stride_bytes, uint64_t rows_cap);
int mma_restore_export(mma_ctx *ctx, uint32_t sample);

/* The same merge for the contexts of ONE process (one per GPU), done by the library: end-of-file flush of every shard, one
 * ncclAllGather of the live rows over NVLink (single-process clique, ncclCommInitAll, kept for the life of the process), import
 * on every GPU.  Afterwards mma_finish_sample on any of the contexts returns the sum TableCount::addCounter would form
 * (mmannot.cpp:1861-1876) had one Counter seen all the shards.  libnccl.so.2 is opened at run time, on the first call with
 * n_ctx > 1 (MMA_ERR_STATE when it cannot be found).  The contexts must have been created with the same parameters; errors are
 * reported on ctxs[0]. */
int mma_allreduce(mma_ctx *const *ctxs, uint32_t n_ctx, uint32_t sample);

int mma_sync(mma_ctx *ctx);
void *mma_stream(mma_ctx *ctx); /* cudaStream_t of the compute stream */

int mma_timing_enable(mma_ctx *ctx, int on);
int mma_timing_reset(mma_ctx *ctx);
int mma_timing_get(mma_ctx *ctx, mma_timing *out); /* synchronises */

/* Size in bytes of the device index built by mma_load_features (0 before), and its number of segments. */
uint64_t mma_index_bytes(const mma_ctx *ctx);
uint64_t mma_index_segments(const mma_ctx *ctx);

/* Bytes mma_finish_sample copies back from the device per sample (table + control block). */
uint64_t mma_readback_bytes(const mma_ctx *ctx);

const char *mma_version(void);
/* Name of the kernel that dominates a batch (for profiles / the roofline line of bench.py). */
const char *mma_dominant_kernel(void);
/* Name of the batch kernel this context launches for its annotation and options ("" before mma_load_features): k_batch_lean with
 * bin entries (annotations up to ~160 Mb), k_batch_fast with the coarser position map, k_batch otherwise. */
const char *mma_batch_kernel(const mma_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MMANNOT_B200_H */
