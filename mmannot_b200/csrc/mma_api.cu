// C ABI of the device hot path (include/mmannot_b200.h): context, streams, staging, launches.
// No CPU fallback lives here: every entry point either drives the sm_100a kernels of
// mma_device.cuh or fails with an error code.
#include "mmannot_b200.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>  // types and prototypes only: the library is opened at run time, by mma_allreduce alone

#include <algorithm>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mma_device.cuh"
#include "mma_batch_fast.cuh"
#include "mma_batch_lean.cuh"
#include "mma_bam.cuh"

using namespace mma;

namespace {

thread_local std::string g_createError;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return ctx->fail(MMA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// bin entries (FastView::ent) are built for annotations of up to this many 64-position bins (83 MB of entries: they have to
// stay L2-resident next to the segment records)
#define MMA_MAX_BIN_ENTRIES 2600000ull

enum TimeCat { TC_INDEX = 0, TC_BATCH, TC_CLOSE, TC_FINISH, TC_N };

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Staging {  // one batch resident on the device
  DevBuf start, end, meta, nh, key;
  DevBuf packed, runKey, tileBase, escIndex, escEnd, escNh;  // compact transfer format, expanded into the five arrays above
  cudaEvent_t copied = nullptr, done = nullptr;
};

struct Sample {
  SampleCtl *ctl = nullptr;
  DevBuf tableKeys, tableVals;
  DevBuf slowKey, slowOrd, slowMask, slowNh, openKeys, openSeq;
  u32 slowCap = 0, openCap = 0;
  // host-side bound on the number of deferred records (see ensureDeferred)
  u32 *countRing = nullptr;  // pinned, 4 entries of slowCount followed by 4 entries of walkCount
  cudaEvent_t ringEv[4] = {nullptr, nullptr, nullptr, nullptr};
  uint64_t ringCum[4] = {0, 0, 0, 0};
  bool ringUsed[4] = {false, false, false, false};
  uint64_t cumHits = 0, seq = 0;
  uint64_t knownCount = 0, knownCum = 0;
  bool touched = false;
  bool openMaybeUsed = false;  // the open-key set may hold keys (unknown until the control block is read back)
  bool deferAll = false;       // this sample's multi-mapping records all go to the deferred list (k_batch_lean RUNS = 2); sticky until reset
  // results
  std::vector<uint64_t> rowMask, rowCount;
  std::vector<uint32_t> rowNh;
};

}  // namespace

struct mma_ctx {
  mma_params params;
  std::vector<uint16_t> elemLine;
  std::vector<uint8_t> elemStrand, elemVic;
  Rules rules;
  bool wideMask = false;
  int device = 0;
  cudaStream_t sc = nullptr, sh = nullptr;
  Staging stage[2];
  uint64_t submitSeq = 0;
  u32 tableCap = 0;
  // index
  bool haveIndex = false;
  DevBuf feat, chrInfo, bins, spanIdx, dElemLine, dElemStrand, dElemVic;
  DevBuf fastBin, fastSeg, fastTie, fastChrInfo, fastEnt, fastRank, fastDict;
  IndexView index;
  FastView fast;
  uint64_t nSegments = 0;
  uint64_t indexBytes = 0;
  std::vector<Sample> samples;
  std::string error;
  // timing
  bool timing = false;
  struct Span { int cat; cudaEvent_t a, b; };
  std::vector<Span> spans;
  struct BamSpan { cudaEvent_t e[5]; };  // inflate e0..e1, record walk + scan e1..e2, (host: the next chunk is read) parse e3..e4
  std::vector<BamSpan> bamSpans;
  std::vector<cudaEvent_t> eventPool;
  double ms[TC_N] = {0, 0, 0, 0};
  uint64_t launches = 0, hitsSubmitted = 0, batches = 0;
  int nSM = 148;
  u32 maxGrid = 0;           // MMANNOT_B200_MAX_GRID=n: cap on k_batch blocks, so that small test inputs still give every warp a multi-tile chunk (testing only)
  size_t randDrawsUsed = 0;  // -y random: rand() draws consumed by the samples finished so far
  DevBuf rndKeys, rndVals;   // -y random over several input files: what the reference's seen / chosenId / numberSeen keep (mm:1742-1747)
  u32 rndCap = 0;
  uint64_t rndNames = 0;     // upper bound of the names held
  int forceGroups = -1;      // MMANNOT_B200_GROUPS=0/1: pin the variant (testing only)
  bool preferDefer = false;  // the last sample / batch left most multi-mapping reads unfinished: start the next sample in DEFER mode
  int forceDefer = -1;       // MMANNOT_B200_DEFER=0/1: pin it (testing only)
  bool useGroups = false;    // k_batch_fast variant for runs of k x NH records, chosen from the walk counters of earlier batches
  int carveout = -1;         // MMANNOT_B200_CARVEOUT=percent: shared-memory carveout of k_batch_lean (tuning only)
  bool legacyBatch = false;  // MMANNOT_B200_LEGACY_BATCH=1: A/B runs of the general k_batch against k_batch_fast (tuning only)
  std::vector<uint32_t> intervalIds;  // result of the last mma_annotate_intervals
  u64 *hostTable = nullptr;  // pinned: [TableDump | rows] of the sample being read back
  DevBuf dumpBuf;            // device side of the same
  DevBuf gatherBuf;          // mma_allreduce: the dumps of all the contexts of the group
  DevBuf defPermA, defPermB, defKeyA, defKeyB, defTmp, defOrd, defMask, defNh, defMax;  // end-of-sample pass over the deferred records (kept: cudaMalloc / cudaFree per sample cost more than the pass)
  // BAM decode on the device (mma_bam.cuh)
  DevBuf bamComp[2], bamOut, bamMemberOff, bamOutOff, bamCount, bamHitOff, bamRefToChr, bamRefFirst, bamFlags;
  DevBuf bamStart, bamEnd, bamMeta, bamNh, bamKey;
  cudaEvent_t bamCopied[2] = {nullptr, nullptr};
  uint64_t bamChunks = 0, bamOrdinal = 0, bamLastHits = 0, bamStaged = 0;
  cudaEvent_t bamStageEv = nullptr;
  u32 bamNRef = 0, bamStrandedness = 1;
  double msBam[3] = {0, 0, 0};  // inflate, count + scan, parse
  struct BamPending {        // the chunk between mma_submit_bam_start and mma_submit_bam_finish
    bool active = false, empty = false;
    u32 sample = 0;
    BamView v{};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  } bamPend;
  u32 bamInflateBlocks = 0;
  u32 *bamPendHost = nullptr;  // page-locked: {flags, hits} of the chunk in flight
  cudaEvent_t bamPendEv = nullptr;
  DevBuf walkMap;            // k_batch_lean -> k_batch_walk: one bit per hit of a launch (zero between launches: the walk clears what it reads)
  DevBuf exportBuf;          // mma_export_table_async: this context's own dump, kept for mma_restore_export (dumpBuf is rewritten by every finish)

  int fail(int code, const std::string &msg) {
    error = msg;
    return code;
  }
  cudaEvent_t getEvent() {
    if (!eventPool.empty()) { cudaEvent_t e = eventPool.back(); eventPool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  struct Timed {
    mma_ctx *c; int cat; cudaEvent_t a = nullptr;
    Timed(mma_ctx *ctx, int category) : c(ctx), cat(category) {
      if (c->timing) { a = c->getEvent(); cudaEventRecord(a, c->sc); }
    }
    ~Timed() {
      if (a) { cudaEvent_t b = c->getEvent(); cudaEventRecord(b, c->sc); c->spans.push_back(Span{cat, a, b}); }
    }
  };
  void collectTiming() {
    for (Span &s : spans) {
      float t = 0;
      if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) ms[s.cat] += t;
      eventPool.push_back(s.a); eventPool.push_back(s.b);
    }
    spans.clear();
    for (BamSpan &b : bamSpans) {
      for (int k = 0; k < 3; ++k) {
        float t = 0;
        if (cudaEventElapsedTime(&t, b.e[k == 2 ? 3 : k], b.e[k == 2 ? 4 : k + 1]) == cudaSuccess) msBam[k] += t;
      }
      for (int k = 0; k < 5; ++k) eventPool.push_back(b.e[k]);
    }
    bamSpans.clear();
  }
};

namespace {

inline u32 gridFor(uint64_t n, u32 threads) { return (u32)((n + threads - 1) / threads); }

TableView tableView(const DevBuf &k, const DevBuf &v, u32 cap, SampleCtl *ctl) {
  TableView t;
  t.keys = k.as<u64>(); t.vals = v.as<u64>(); t.capMask = cap - 1; t.overflow = &ctl->overflow;
  return t;
}
SlowView slowView(const Sample &s) {
  SlowView v;
  v.key = s.slowKey.as<u64>(); v.ord = s.slowOrd.as<u64>(); v.mask = s.slowMask.as<u64>(); v.nh = s.slowNh.as<u32>(); v.cap = s.slowCap;
  return v;
}
KeySetView openView(const Sample &s) {
  KeySetView v;
  v.keys = s.openKeys.as<u64>(); v.seq = s.openSeq.as<u32>(); v.capMask = s.openCap ? s.openCap - 1 : 0;
  return v;
}

int initSample(mma_ctx *ctx, Sample &s) {
  if (s.ctl) return MMA_OK;
  CK(cudaMalloc(&s.ctl, sizeof(SampleCtl)));
  CK(cudaMemsetAsync(s.ctl, 0, sizeof(SampleCtl), ctx->sc));
  const size_t tb = (size_t)ctx->tableCap * sizeof(u64);
  CK(s.tableKeys.ensure(tb)); CK(s.tableVals.ensure(tb));
  CK(cudaMemsetAsync(s.tableKeys.p, 0, tb, ctx->sc)); CK(cudaMemsetAsync(s.tableVals.p, 0, tb, ctx->sc));
  CK(cudaHostAlloc(&s.countRing, 8 * sizeof(u32), cudaHostAllocDefault));
  for (int i = 0; i < 4; ++i) CK(cudaEventCreateWithFlags(&s.ringEv[i], cudaEventDisableTiming));
  return MMA_OK;
}

__global__ void k_and_u32(u32 *p, u32 mask) { *p &= mask; }

__global__ void k_keyset_rehash(KeySetView from, KeySetView to) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > from.capMask) return;
  const u64 k = from.keys[i];
  if (k == KEY_EMPTY) return;
  u32 slot = (u32)mix64(k) & to.capMask;
  for (;;) {
    const u64 o = atomicCAS(&to.keys[slot], KEY_EMPTY, k);
    if (o == KEY_EMPTY || o == k) { to.seq[slot] = from.seq[i]; return; }
    slot = (slot + 1) & to.capMask;
  }
}

// Deferred list / open-key set sizing.  The device appends without asking; the host keeps an
// upper bound on the list length (last count read back asynchronously + hits submitted since,
// plus one stand-in record per batch for a carried read) and only synchronises when that bound
// says the next batch might not fit.
int ensureDeferred(mma_ctx *ctx, Sample &s, uint64_t n) {
  const bool needs = (ctx->rules.strategy == MMA_STRATEGY_DEFAULT || ctx->rules.strategy == MMA_STRATEGY_RANDOM);
  if (!needs) return MMA_OK;
  n += 2;
  // freshest completed read-back
  for (int k = 0; k < 4; ++k) {
    const int slot = (int)((s.seq + 3 - k) & 3);  // newest first
    if (!s.ringUsed[slot]) continue;
    if (cudaEventQuery(s.ringEv[slot]) == cudaSuccess) {
      if (s.ringCum[slot] >= s.knownCum) { s.knownCount = s.countRing[slot]; s.knownCum = s.ringCum[slot]; }
      // many serial walks (or, in the GROUPS variant, many groups beyond the first of their run) per hit so far: paired-end
      // like input, keep / switch to the GROUPS variant; hardly any: the plain variant (hysteresis in between)
      const uint64_t walks = s.countRing[4 + slot], hitsSoFar = std::max<uint64_t>(s.ringCum[slot], 1);
      if (ctx->forceGroups < 0) {
        if (walks * 32 > hitsSoFar) ctx->useGroups = true;
        else if (walks * 256 < hitsSoFar) ctx->useGroups = false;
      }
      // many deferred records per hit: the records of a read are not adjacent (coordinate-sorted file) -- stop running the
      // countdown in the batch kernel and defer every multi-mapping record of the rest of the sample
      if (ctx->forceDefer < 0 && !s.deferAll && (uint64_t)s.countRing[slot] * 8 > hitsSoFar) { s.deferAll = true; ctx->preferDefer = true; }
      break;
    }
  }
  uint64_t upper = s.knownCount + (s.cumHits - s.knownCum);
  if (s.slowCap != 0 && upper + n <= s.slowCap) return MMA_OK;
  uint64_t actual = 0;
  if (s.slowCap != 0) {
    CK(cudaStreamSynchronize(ctx->sc));
    u32 c = 0;
    CK(cudaMemcpy(&c, &s.ctl->slowCount, sizeof(u32), cudaMemcpyDeviceToHost));
    actual = c;
    s.knownCount = actual; s.knownCum = s.cumHits;
    if (actual + n <= s.slowCap) return MMA_OK;
  }
  uint64_t want = std::max<uint64_t>(std::max<uint64_t>(2ull * ctx->params.max_batch_hits, 1u << 16), 2 * (actual + n));
  if (want > 0x7FFFFFF0ull) return ctx->fail(MMA_ERR_CAPACITY, "deferred read list would exceed 2^31 records");  // (the radix sort takes int counts)
  u32 newCap = (u32)want;
  DevBuf nk, no, nm, nn;
  CK(nk.ensure((size_t)newCap * 8)); CK(no.ensure((size_t)newCap * 8)); CK(nm.ensure((size_t)newCap * 8)); CK(nn.ensure((size_t)newCap * 4));
  if (actual) {
    CK(cudaMemcpyAsync(nk.p, s.slowKey.p, actual * 8, cudaMemcpyDeviceToDevice, ctx->sc));
    CK(cudaMemcpyAsync(no.p, s.slowOrd.p, actual * 8, cudaMemcpyDeviceToDevice, ctx->sc));
    CK(cudaMemcpyAsync(nm.p, s.slowMask.p, actual * 8, cudaMemcpyDeviceToDevice, ctx->sc));
    CK(cudaMemcpyAsync(nn.p, s.slowNh.p, actual * 4, cudaMemcpyDeviceToDevice, ctx->sc));
  }
  u32 newOpen = 1;
  while (newOpen < 2ull * newCap && newOpen < 0x80000000u) newOpen <<= 1;
  DevBuf nopen, nseq;
  CK(nopen.ensure((size_t)newOpen * 8)); CK(nseq.ensure((size_t)newOpen * 4));
  k_fill_u64<<<gridFor(newOpen, 256), 256, 0, ctx->sc>>>(nopen.as<u64>(), KEY_EMPTY, newOpen);
  k_fill_u32<<<gridFor(newOpen, 256), 256, 0, ctx->sc>>>(nseq.as<u32>(), 0xFFFFFFFFu, newOpen);
  ctx->launches += 2;
  if (s.openCap) {
    KeySetView from = openView(s), to;
    to.keys = nopen.as<u64>(); to.seq = nseq.as<u32>(); to.capMask = newOpen - 1;
    k_keyset_rehash<<<gridFor(s.openCap, 256), 256, 0, ctx->sc>>>(from, to);
    ctx->launches++;
  }
  CK(cudaStreamSynchronize(ctx->sc));
  s.slowKey.release(); s.slowOrd.release(); s.slowMask.release(); s.slowNh.release(); s.openKeys.release(); s.openSeq.release();
  s.slowKey = nk; s.slowOrd = no; s.slowMask = nm; s.slowNh = nn; s.openKeys = nopen; s.openSeq = nseq;
  s.slowCap = newCap; s.openCap = newOpen;
  return MMA_OK;
}

// The batch kernels keep per-thread counters in 16-bit fields (and 16-bit histogram columns): a warp must not see more than
// 2^14 tiles of 128 hits in one launch.  With the usual grids that is far away (2^32 hits per batch); a capped grid
// (MMANNOT_B200_MAX_GRID, small GPUs) reaches it earlier, so longer batches are cut into several launches -- a cut is just a
// batch border, which the kernels handle anyway.
#define MMA_MAX_TILES_PER_WARP 16000ull

template <int MODE, int STRAT, bool FAST, typename MaskT>
void launchBatchKernels(mma_ctx *ctx, Sample &s, const HitView &hAll) {
  {
    const u32 minGrid = ctx->maxGrid ? std::min<u32>(ctx->maxGrid, (u32)ctx->nSM) : (u32)ctx->nSM;
    const uint64_t maxHits = MMA_MAX_TILES_PER_WARP * WT_HITS * (uint64_t)minGrid * 8ull;
    if (hAll.n > maxHits) {
      const u32 n1 = (u32)(maxHits);  // (a multiple of 128: the second part keeps the alignment of the arrays)
      HitView a = hAll, b = hAll;
      a.n = n1;
      b.n = hAll.n - n1; b.start += n1; b.end += n1; b.meta += n1; b.nh += n1; b.key += n1;
      launchBatchKernels<MODE, STRAT, FAST, MaskT>(ctx, s, a);
      launchBatchKernels<MODE, STRAT, FAST, MaskT>(ctx, s, b);
      return;
    }
  }
  const HitView &h = hAll;
  const Rules &r = ctx->rules;
  TableView table = tableView(s.tableKeys, s.tableVals, ctx->tableCap, s.ctl);
  SlowView slow = slowView(s);
  KeySetView open = openView(s);
  // one contiguous chunk of 128-hit warp tiles per warp; enough warps to fill the GPU, chunks of >= 8 tiles when possible
  const u32 nWT = (h.n + WT_HITS - 1) / WT_HITS;
  constexpr bool useFast = FAST && sizeof(MaskT) == 4 && STRAT != 2;  // k_batch_lean / k_batch_fast
  bool launched = false;
  if constexpr (useFast) if (!ctx->legacyBatch && ctx->fast.nChr <= CHR_SMEM) {
    // runs of k x NH records (paired-end data): the GROUPS variant once a batch has shown many of them (see afterBatch)
    const bool groups = STRAT == 0 && ctx->useGroups;
    const bool defer = STRAT == 0 && s.deferAll;
    if (ctx->fast.ent && h.vec) {  // bin entries: the hit arrays go through the TMA ring (needs 16-byte aligned arrays)
      launched = true;
      u32 grid = std::max<u32>(1u, std::min<u32>((nWT + LEAN_WARPS - 1) / LEAN_WARPS, (u32)ctx->nSM * MMA_LEAN_BLOCKS_PER_SM));
      if (ctx->maxGrid) grid = std::min(grid, ctx->maxGrid);
      const size_t smem = leanSmemBytes<STRAT != 3>(ctx->fast.nDict, ctx->fast.nChr, r.nElements);
      auto kernel = defer ? k_batch_lean<MODE, STRAT, (STRAT == 0 ? 2 : 0)> : groups ? k_batch_lean<MODE, STRAT, (STRAT == 0 ? 1 : 0)> : k_batch_lean<MODE, STRAT, 0>;
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // (per device: cheap enough per launch)
      {  // shared memory for the blocks one SM can hold, the rest of the 228 KB stays L1 (the bin entries of neighbouring hits hit there)
        const size_t perSM = (size_t)MMA_LEAN_BLOCKS_PER_SM * (smem + 1024);
        int pct = (int)std::min<size_t>(100, (perSM * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (ctx->carveout >= 0) pct = ctx->carveout;
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      }
      u32 *walkMap = ctx->walkMap.as<u32>();  // (sized by launchBatch)
      {
        mma_ctx::Timed t(ctx, TC_BATCH);
        kernel<<<grid, LEAN_THREADS, smem, ctx->sc>>>(ctx->index, ctx->fast, h, r, table, s.ctl, slow, walkMap);
      }
      if (STRAT == 0) {  // the runs the scan left: one thread per marked run (usually none), and the read carried into the batch
        mma_ctx::Timed t(ctx, TC_CLOSE);
        const u32 gridW = defer ? 1u : std::max<u32>(1u, std::min<u32>(gridFor((h.n + 127) / 128, 256), (u32)ctx->nSM * 8u));  // a lane reads 128 bits per round
        k_batch_walk<MODE><<<gridW, 256, 0, ctx->sc>>>(ctx->index, ctx->fast, h, r, table, s.ctl, slow, open, walkMap, defer ? 1 : 0);
        ctx->launches += 1;
      }
    } else if (ctx->fast.bm) {
      launched = true;
      u32 grid = std::max<u32>(1u, std::min<u32>((nWT + FAST_WARPS - 1) / FAST_WARPS, (u32)ctx->nSM * MMA_FAST_BLOCKS_PER_SM));
      if (ctx->maxGrid) grid = std::min(grid, ctx->maxGrid);
      u32 *walkMap = ctx->walkMap.as<u32>();  // (sized by launchBatch)
      {
        mma_ctx::Timed t(ctx, TC_BATCH);
        if (groups) k_batch_fast<MODE, STRAT, (STRAT == 0)><<<grid, FAST_THREADS, 0, ctx->sc>>>(ctx->index, ctx->fast, h, r, table, s.ctl, slow, walkMap);
        else k_batch_fast<MODE, STRAT, false><<<grid, FAST_THREADS, 0, ctx->sc>>>(ctx->index, ctx->fast, h, r, table, s.ctl, slow, walkMap);
      }
      if (STRAT == 0) {
        mma_ctx::Timed t(ctx, TC_CLOSE);
        const u32 gridW = std::max<u32>(1u, std::min<u32>(gridFor((h.n + 127) / 128, 256), (u32)ctx->nSM * 8u));
        k_batch_walk<MODE><<<gridW, 256, 0, ctx->sc>>>(ctx->index, ctx->fast, h, r, table, s.ctl, slow, open, walkMap, 0);
        ctx->launches += 1;
      }
    }
  }
  if (!launched) {
    const u32 perSM = (sizeof(MaskT) == 4) ? MMA_BLOCKS_PER_SM : 2;
    u32 grid = std::max<u32>(1u, std::min<u32>((nWT + BATCH_WARPS - 1) / BATCH_WARPS, (u32)ctx->nSM * perSM));
    if (ctx->maxGrid) grid = std::min(grid, ctx->maxGrid);
    mma_ctx::Timed t(ctx, TC_BATCH);
    k_batch<MODE, STRAT, FAST, MaskT><<<grid, BATCH_THREADS, 0, ctx->sc>>>(ctx->index, ctx->fast, h, r, table, s.ctl, slow, open);
  }
  {
    mma_ctx::Timed t(ctx, TC_CLOSE);
    const u32 gridC = (STRAT == 0) ? std::min<u32>(gridFor(h.n, 256), (u32)ctx->nSM) : 1u;
    k_batch_close<MODE, FAST><<<gridC, 256, 0, ctx->sc>>>(ctx->index, ctx->fast, h, r, table, s.ctl, slow, open);
  }
  ctx->launches += 2;
  ctx->batches++;
}

template <int MODE, bool FAST, typename MaskT>
void launchBatchStrategy(mma_ctx *ctx, Sample &s, const HitView &h) {
  switch (ctx->rules.strategy) {
    case MMA_STRATEGY_DEFAULT: launchBatchKernels<MODE, 0, FAST, MaskT>(ctx, s, h); break;
    case MMA_STRATEGY_UNIQUE: launchBatchKernels<MODE, 1, FAST, MaskT>(ctx, s, h); break;
    case MMA_STRATEGY_RANDOM: launchBatchKernels<MODE, 2, FAST, MaskT>(ctx, s, h); break;
    default: launchBatchKernels<MODE, 3, FAST, MaskT>(ctx, s, h); break;
  }
}

template <int MODE>
void launchBatchMode(mma_ctx *ctx, Sample &s, const HitView &h) {
  if (ctx->wideMask) launchBatchStrategy<MODE, false, u64>(ctx, s, h);
  else if (ctx->fast.enabled) launchBatchStrategy<MODE, true, u32>(ctx, s, h);
  else launchBatchStrategy<MODE, false, u32>(ctx, s, h);
}

int launchBatch(mma_ctx *ctx, Sample &s, const HitView &h) {
  if (ctx->rules.strategy == MMA_STRATEGY_DEFAULT && ctx->fast.enabled) {  // k_batch_lean / k_batch_fast mark the runs they leave to k_batch_walk: a bit per hit
    const size_t need = ((size_t)h.n + 127) / 128 * 16 + 16;
    if (need > ctx->walkMap.bytes) {
      const size_t want = std::max<size_t>(need, ((size_t)ctx->params.max_batch_hits + 127) / 128 * 16 + 16);
      if (ctx->walkMap.ensure(want) != cudaSuccess) return ctx->fail(MMA_ERR_CUDA, "out of device memory (walk map)");
      CK(cudaMemsetAsync(ctx->walkMap.p, 0, ctx->walkMap.bytes, ctx->sc));
    }
  }
  if (ctx->rules.mode == 0) launchBatchMode<0>(ctx, s, h);
  else if (ctx->rules.mode == 1) launchBatchMode<1>(ctx, s, h);
  else launchBatchMode<2>(ctx, s, h);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ctx->fail(MMA_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  return MMA_OK;
}

int afterBatch(mma_ctx *ctx, Sample &s, uint64_t n) {
  s.cumHits += n + 2;
  s.touched = true;
  s.openMaybeUsed = true;
  ctx->hitsSubmitted += n;
  const int slot = (int)(s.seq & 3);
  CK(cudaMemcpyAsync(&s.countRing[slot], &s.ctl->slowCount, sizeof(u32), cudaMemcpyDeviceToHost, ctx->sc));
  CK(cudaMemcpyAsync(&s.countRing[4 + slot], &s.ctl->walkCount, sizeof(u32), cudaMemcpyDeviceToHost, ctx->sc));
  CK(cudaEventRecord(s.ringEv[slot], ctx->sc));
  s.ringCum[slot] = s.cumHits;
  s.ringUsed[slot] = true;
  s.seq++;
  return MMA_OK;
}

int checkSubmit(mma_ctx *ctx, uint32_t sample, const mma_hit_batch *b) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!b) return ctx->fail(MMA_ERR_INVALID, "null batch");
  if (!ctx->haveIndex) return ctx->fail(MMA_ERR_STATE, "mma_load_features must be called before hits are submitted");
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  if (b->n > ctx->params.max_batch_hits) return ctx->fail(MMA_ERR_INVALID, "batch larger than max_batch_hits");
  if (b->n && (!b->start || !b->end || !b->meta || !b->nh || !b->read_key)) return ctx->fail(MMA_ERR_INVALID, "null hit array");
  return MMA_OK;
}

}  // namespace

extern "C" {

const char *mma_version(void) { return "mmannot_b200 0.1 (sm_100a)"; }

const char *mma_last_error(const mma_ctx *ctx) { return ctx ? ctx->error.c_str() : g_createError.c_str(); }

int mma_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int mma_warmup(int device) {
  if (cudaSetDevice(device) != cudaSuccess) return MMA_ERR_CUDA;
  return cudaFree(nullptr) == cudaSuccess ? MMA_OK : MMA_ERR_CUDA;
}

int mma_create(mma_ctx **out, const mma_params *p) {
  if (!out || !p) { g_createError = "null argument"; return MMA_ERR_INVALID; }
  *out = nullptr;
  if (p->n_elements == 0 || p->n_elements > MMA_MAX_ELEMENTS || !p->elem_line || !p->elem_strand || !p->elem_vicinity) {
    g_createError = "n_elements must be in 1..64 and the element tables must be given";
    return MMA_ERR_INVALID;
  }
  if (p->strategy < 0 || p->strategy > 3) { g_createError = "unknown strategy"; return MMA_ERR_INVALID; }
  if (p->strategy == MMA_STRATEGY_RATIO && p->n_elements > NH_SHIFT) {
    g_createError = "-y ratio supports at most 40 elements in the Order section";
    return MMA_ERR_INVALID;
  }
  if (p->n_samples == 0 || p->max_batch_hits == 0) { g_createError = "n_samples and max_batch_hits must be positive"; return MMA_ERR_INVALID; }
  for (uint32_t i = 0; i < p->n_elements; ++i)
    if (p->elem_strand[i] > 2 || p->elem_vicinity[i] > 2) { g_createError = "bad element table entry"; return MMA_ERR_INVALID; }
  int nDev = 0;
  cudaError_t e = cudaGetDeviceCount(&nDev);
  if (e != cudaSuccess || nDev == 0) {
    g_createError = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); mmannot_b200 has no CPU fallback";
    return MMA_ERR_NO_DEVICE;
  }
  if (p->device < 0 || p->device >= nDev) { g_createError = "device ordinal out of range"; return MMA_ERR_INVALID; }
  mma_ctx *ctx = new mma_ctx();
  ctx->params = *p;
  ctx->device = p->device;
  ctx->elemLine.assign(p->elem_line, p->elem_line + p->n_elements);
  ctx->elemStrand.assign(p->elem_strand, p->elem_strand + p->n_elements);
  ctx->elemVic.assign(p->elem_vicinity, p->elem_vicinity + p->n_elements);
  ctx->params.elem_line = ctx->elemLine.data();
  ctx->params.elem_strand = ctx->elemStrand.data();
  ctx->params.elem_vicinity = ctx->elemVic.data();
  Rules &r = ctx->rules;
  r.strategy = p->strategy;
  r.overlap = p->overlap;
  r.mode = (p->overlap < 0.0f) ? 0 : (p->overlap < 1.0f) ? 1 : 2;  // mm:1974-1976
  r.rescueThreshold = p->rescue_threshold;
  r.rescue = (p->read_stats != 0 && p->rescue_threshold < 1.0f) ? 1 : 0;  // mm:491, 2025
  r.nElements = p->n_elements;
  ctx->wideMask = p->n_elements > 30;  // bits 30 and 31 of a 32-bit element set serve as flags (segment table, run scan)
  ctx->tableCap = 1u << (p->table_log2 ? std::min<uint32_t>(std::max<uint32_t>(p->table_log2, 8), 26) : 16);
  auto bail = [&](const char *what, cudaError_t err) {
    g_createError = std::string(what) + ": " + cudaGetErrorString(err);
    mma_destroy(ctx);
    return MMA_ERR_CUDA;
  };
  if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaDeviceGetAttribute(&ctx->nSM, cudaDevAttrMultiProcessorCount, ctx->device)) != cudaSuccess) return bail("cudaDeviceGetAttribute", e);
  ctx->fast = FastView();
  if ((e = cudaStreamCreateWithFlags(&ctx->sc, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->sh, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (int k = 0; k < 2; ++k) {
    if ((e = cudaEventCreateWithFlags(&ctx->stage[k].copied, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->stage[k].done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  }
  { const char *lg = getenv("MMANNOT_B200_LEGACY_BATCH"); ctx->legacyBatch = lg && lg[0] == '1';
    const char *fg = getenv("MMANNOT_B200_GROUPS"); if (fg && (fg[0] == '0' || fg[0] == '1')) { ctx->forceGroups = fg[0] - '0'; ctx->useGroups = fg[0] == '1'; }
    const char *fd = getenv("MMANNOT_B200_DEFER"); if (fd && (fd[0] == '0' || fd[0] == '1')) { ctx->forceDefer = fd[0] - '0'; ctx->preferDefer = fd[0] == '1'; }
    const char *cv = getenv("MMANNOT_B200_CARVEOUT"); if (cv) ctx->carveout = atoi(cv);
    const char *mg = getenv("MMANNOT_B200_MAX_GRID"); ctx->maxGrid = mg ? (u32)std::max(0, atoi(mg)) : 0u; }
  ctx->samples.resize(p->n_samples);
  *out = ctx;
  return MMA_OK;
}

void mma_destroy(mma_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->sc) cudaStreamSynchronize(ctx->sc);
  if (ctx->sh) cudaStreamSynchronize(ctx->sh);
  for (Sample &s : ctx->samples) {
    if (s.ctl) cudaFree(s.ctl);
    s.tableKeys.release(); s.tableVals.release();
    s.slowKey.release(); s.slowOrd.release(); s.slowMask.release(); s.slowNh.release(); s.openKeys.release(); s.openSeq.release();
    if (s.countRing) cudaFreeHost(s.countRing);
    for (int i = 0; i < 4; ++i) if (s.ringEv[i]) cudaEventDestroy(s.ringEv[i]);
  }
  for (int k = 0; k < 2; ++k) {
    Staging &g = ctx->stage[k];
    g.start.release(); g.end.release(); g.meta.release(); g.nh.release(); g.key.release();
    g.packed.release(); g.runKey.release(); g.tileBase.release(); g.escIndex.release(); g.escEnd.release(); g.escNh.release();
    if (g.copied) cudaEventDestroy(g.copied);
    if (g.done) cudaEventDestroy(g.done);
  }
  ctx->feat.release(); ctx->chrInfo.release(); ctx->bins.release(); ctx->spanIdx.release();
  ctx->fastBin.release(); ctx->fastSeg.release(); ctx->fastTie.release(); ctx->fastChrInfo.release(); ctx->fastEnt.release(); ctx->fastRank.release(); ctx->fastDict.release();
  ctx->dElemLine.release(); ctx->dElemStrand.release(); ctx->dElemVic.release();
  ctx->collectTiming();
  if (ctx->hostTable) cudaFreeHost(ctx->hostTable);
  ctx->dumpBuf.release();
  ctx->gatherBuf.release();
  ctx->exportBuf.release();
  ctx->walkMap.release();
  ctx->rndKeys.release(); ctx->rndVals.release();
  for (int k = 0; k < 2; ++k) { ctx->bamComp[k].release(); if (ctx->bamCopied[k]) cudaEventDestroy(ctx->bamCopied[k]); }
  if (ctx->bamStageEv) cudaEventDestroy(ctx->bamStageEv);
  if (ctx->bamPendEv) cudaEventDestroy(ctx->bamPendEv);
  if (ctx->bamPendHost) cudaFreeHost(ctx->bamPendHost);
  ctx->bamOut.release(); ctx->bamMemberOff.release(); ctx->bamOutOff.release(); ctx->bamCount.release(); ctx->bamHitOff.release();
  ctx->bamRefToChr.release(); ctx->bamRefFirst.release(); ctx->bamFlags.release();
  ctx->bamStart.release(); ctx->bamEnd.release(); ctx->bamMeta.release(); ctx->bamNh.release(); ctx->bamKey.release();
  ctx->defPermA.release(); ctx->defPermB.release(); ctx->defKeyA.release(); ctx->defKeyB.release(); ctx->defTmp.release();
  ctx->defOrd.release(); ctx->defMask.release(); ctx->defNh.release(); ctx->defMax.release();
  for (cudaEvent_t e : ctx->eventPool) cudaEventDestroy(e);
  if (ctx->sc) cudaStreamDestroy(ctx->sc);
  if (ctx->sh) cudaStreamDestroy(ctx->sh);
  delete ctx;
}

void *mma_alloc_pinned(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void mma_free_pinned(void *p) {
  if (p) cudaFreeHost(p);
}

int mma_load_features(mma_ctx *ctx, const mma_features *f) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!f || f->n == 0 || !f->chr || !f->start || !f->end || !f->type || !f->strand) return ctx->fail(MMA_ERR_INVALID, "empty or null feature buffer");
  if (f->n_chr == 0 || f->n_chr >= MMA_HIT_CHR_NONE) return ctx->fail(MMA_ERR_INVALID, "n_chr out of range");
  CK(cudaSetDevice(ctx->device));
  const uint32_t n = f->n, nChr = f->n_chr;
  // host pass: validate the order, find chromosome ranges and extents
  std::vector<u32> chrStart(nChr + 1, 0);
  std::vector<uint64_t> extent(nChr, 0);
  for (uint32_t i = 0; i < n; ++i) {
    if (f->chr[i] >= nChr) return ctx->fail(MMA_ERR_INVALID, "feature chromosome id >= n_chr");
    if (f->type[i] >= ctx->params.n_elements) return ctx->fail(MMA_ERR_INVALID, "feature type >= n_elements");
    if (f->start[i] >= 0xFFFFFFF0u || f->end[i] >= 0xFFFFFFF0u) return ctx->fail(MMA_ERR_INVALID, "feature coordinates >= 0xFFFFFFF0 are reserved");
    if (i > 0 && (f->chr[i] < f->chr[i - 1] || (f->chr[i] == f->chr[i - 1] && f->start[i] < f->start[i - 1])))
      return ctx->fail(MMA_ERR_INVALID, "features must be sorted by (chromosome, start)");
    chrStart[f->chr[i] + 1]++;
    extent[f->chr[i]] = std::max<uint64_t>(extent[f->chr[i]], std::max(f->start[i], f->end[i]));
  }
  for (uint32_t c = 0; c < nChr; ++c) chrStart[c + 1] += chrStart[c];
  uint32_t shift = ctx->params.bin_shift;
  if (shift == 0) {
    uint64_t total = 0;
    for (uint32_t c = 0; c < nChr; ++c) total += extent[c];
    shift = 5;
    while (shift < 20 && (total >> shift) > (1ull << 21)) ++shift;
  }
  shift = std::min<uint32_t>(std::max<uint32_t>(shift, 1), 24);
  std::vector<u32> chrBinBase(nChr + 1, 0);
  std::vector<uint2> chrInfo(nChr);
  uint64_t entries = 0;
  for (uint32_t c = 0; c < nChr; ++c) {
    const uint64_t nb = (chrStart[c + 1] > chrStart[c]) ? (extent[c] >> shift) + 2 : 1;
    chrBinBase[c] = (u32)entries;
    chrInfo[c] = make_uint2((u32)entries, (u32)nb);
    entries += nb + 1;
    if (entries > 0x7FFFFFFFull) return ctx->fail(MMA_ERR_INVALID, "position index too large; raise bin_shift");
  }
  chrBinBase[nChr] = (u32)entries;

  DevBuf dChr, dStart, dEnd, dType, dStrand, dChrStart, dChrBinBase, dSpanCount, dSpanOff, dScanTmp, dTotal;
  auto cleanup = [&]() { dChr.release(); dStart.release(); dEnd.release(); dType.release(); dStrand.release(); dChrStart.release(); dChrBinBase.release(); dSpanCount.release(); dSpanOff.release(); dScanTmp.release(); dTotal.release(); };
#define CKL(call)                                                                                  \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) { cleanup(); return ctx->fail(MMA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } \
  } while (0)
  CKL(dChr.ensure((size_t)n * 4)); CKL(dStart.ensure((size_t)n * 4)); CKL(dEnd.ensure((size_t)n * 4));
  CKL(dType.ensure(n)); CKL(dStrand.ensure(n));
  CKL(dChrStart.ensure((size_t)(nChr + 1) * 4)); CKL(dChrBinBase.ensure((size_t)(nChr + 1) * 4));
  CKL(dSpanCount.ensure((size_t)entries * 4)); CKL(dSpanOff.ensure((size_t)entries * 4)); CKL(dTotal.ensure(4));
  size_t scanBytes = 0;
  CKL(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, dSpanCount.as<u32>(), dSpanOff.as<u32>(), (int)entries, ctx->sc));
  CKL(dScanTmp.ensure(scanBytes ? scanBytes : 1));
  CKL(ctx->feat.ensure((size_t)n * sizeof(uint4)));
  CKL(ctx->chrInfo.ensure((size_t)nChr * sizeof(uint2)));
  CKL(ctx->bins.ensure((size_t)entries * sizeof(uint2)));
  CKL(ctx->dElemLine.ensure(ctx->elemLine.size() * 2)); CKL(ctx->dElemStrand.ensure(ctx->elemStrand.size())); CKL(ctx->dElemVic.ensure(ctx->elemVic.size()));
  cudaStream_t st = ctx->sc;
  CKL(cudaMemcpyAsync(dChr.p, f->chr, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(dStart.p, f->start, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(dEnd.p, f->end, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(dType.p, f->type, n, cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(dStrand.p, f->strand, n, cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(dChrStart.p, chrStart.data(), (size_t)(nChr + 1) * 4, cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(dChrBinBase.p, chrBinBase.data(), (size_t)(nChr + 1) * 4, cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(ctx->chrInfo.p, chrInfo.data(), (size_t)nChr * sizeof(uint2), cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(ctx->dElemLine.p, ctx->elemLine.data(), ctx->elemLine.size() * 2, cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(ctx->dElemStrand.p, ctx->elemStrand.data(), ctx->elemStrand.size(), cudaMemcpyHostToDevice, st));
  CKL(cudaMemcpyAsync(ctx->dElemVic.p, ctx->elemVic.data(), ctx->elemVic.size(), cudaMemcpyHostToDevice, st));
  BuildView b;
  b.chr = dChr.as<u32>(); b.start = dStart.as<u32>(); b.end = dEnd.as<u32>();
  b.type = dType.as<uint8_t>(); b.strand = dStrand.as<uint8_t>();
  b.elemLine = ctx->dElemLine.as<uint16_t>(); b.elemStrand = ctx->dElemStrand.as<uint8_t>(); b.elemVic = ctx->dElemVic.as<uint8_t>();
  b.chrStart = dChrStart.as<u32>(); b.chrBinBase = dChrBinBase.as<u32>();
  b.nFeat = n; b.nChr = nChr; b.shift = shift; b.nEntries = (u32)entries;
  u32 total = 0;
  {
    mma_ctx::Timed t(ctx, TC_INDEX);
    k_pack_features<<<gridFor(n, 256), 256, 0, st>>>(b, ctx->feat.as<uint4>());
    k_prefix_max_end<<<nChr, 256, 0, st>>>(b, ctx->feat.as<uint4>());
    k_build_bins<<<gridFor(entries, 128), 128, 0, st>>>(b, ctx->feat.as<uint4>(), ctx->bins.as<uint2>(), dSpanCount.as<u32>(), nullptr, 0);
    CKL(cub::DeviceScan::ExclusiveSum(dScanTmp.p, scanBytes, dSpanCount.as<u32>(), dSpanOff.as<u32>(), (int)entries, st));
    k_set_span_offsets<<<gridFor(entries, 256), 256, 0, st>>>(dSpanOff.as<u32>(), dSpanCount.as<u32>(), ctx->bins.as<uint2>(), (u32)entries, dTotal.as<u32>());
    ctx->launches += 4;
  }
  CKL(cudaMemcpyAsync(&total, dTotal.p, 4, cudaMemcpyDeviceToHost, st));
  CKL(cudaStreamSynchronize(st));
  CKL(ctx->spanIdx.ensure((size_t)std::max<u32>(total, 1) * 4));
  {
    mma_ctx::Timed t(ctx, TC_INDEX);
    k_build_bins<<<gridFor(entries, 128), 128, 0, st>>>(b, ctx->feat.as<uint4>(), ctx->bins.as<uint2>(), dSpanCount.as<u32>(), ctx->spanIdx.as<u32>(), 1);
    ctx->launches++;
  }
  CKL(cudaStreamSynchronize(st));
  CKL(cudaGetLastError());
  // ---- segment answer table (E <= 31 only: bit 31 of an answer word is the "position dependent" flag)
  ctx->index.feat = ctx->feat.as<uint4>();
  ctx->index.chrInfo = ctx->chrInfo.as<uint2>();
  ctx->index.bins = ctx->bins.as<uint2>();
  ctx->index.spanIdx = ctx->spanIdx.as<u32>();
  ctx->index.nChr = nChr;
  ctx->index.shift = shift;
  ctx->fast = FastView();
  ctx->nSegments = 0;
  uint64_t fastBytes = 0;
  if (ctx->params.n_elements <= 30 && ctx->params.fast_bin_shift != MMA_FAST_OFF) {
    const uint64_t nKeys = 2ull * n + nChr;
    DevBuf kA, kB, flag, pos, tmp, segKey, fChrBinBase, dictHash;
    u32 nDictUsed = 2;
    auto cleanupFast = [&]() { kA.release(); kB.release(); flag.release(); pos.release(); tmp.release(); segKey.release(); fChrBinBase.release(); dictHash.release(); };
#define CKS(call)                                                                                  \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) { cleanupFast(); cleanup(); return ctx->fail(MMA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } \
  } while (0)
    CKS(kA.ensure(nKeys * 8)); CKS(kB.ensure(nKeys * 8)); CKS(flag.ensure(nKeys * 4)); CKS(pos.ensure(nKeys * 4));
    size_t tb1 = 0, tb2 = 0;
    CKS(cub::DeviceRadixSort::SortKeys(nullptr, tb1, kA.as<u64>(), kB.as<u64>(), (int)nKeys, 0, 64, st));
    CKS(cub::DeviceScan::ExclusiveSum(nullptr, tb2, flag.as<u32>(), pos.as<u32>(), (int)nKeys, st));
    CKS(tmp.ensure(std::max(tb1, tb2)));
    u32 lastFlag = 0, lastPos = 0;
    {
      mma_ctx::Timed t(ctx, TC_INDEX);
      k_seg_keys<<<gridFor(std::max<u32>(n, nChr), 256), 256, 0, st>>>(b, kA.as<u64>());
      CKS(cub::DeviceRadixSort::SortKeys(tmp.p, tb1, kA.as<u64>(), kB.as<u64>(), (int)nKeys, 0, 64, st));
      k_seg_flag<<<gridFor(nKeys, 256), 256, 0, st>>>(kB.as<u64>(), (u32)nKeys, flag.as<u32>());
      CKS(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, flag.as<u32>(), pos.as<u32>(), (int)nKeys, st));
      ctx->launches += 2;
    }
    CKS(cudaMemcpyAsync(&lastFlag, flag.as<u32>() + (nKeys - 1), 4, cudaMemcpyDeviceToHost, st));
    CKS(cudaMemcpyAsync(&lastPos, pos.as<u32>() + (nKeys - 1), 4, cudaMemcpyDeviceToHost, st));
    CKS(cudaStreamSynchronize(st));
    const u32 nSeg = lastPos + lastFlag;
    // Annotations up to ~160 Mb (auto mode only): BIN ENTRIES, one 32-byte entry per 64 positions that answers the common read
    // with a single gather (k_batch_lean).  Otherwise the position map: 32 granules per entry; one position per granule when the
    // map stays within 2^22 entries (32 MB), coarser granules for larger annotations.
    uint32_t fshift = ctx->params.fast_bin_shift;
    uint64_t totalExtent = 0;
    for (uint32_t c = 0; c < nChr; ++c) totalExtent += extent[c];
    uint64_t maxBins = MMA_MAX_BIN_ENTRIES;
    if (const char *mb = getenv("MMANNOT_B200_MAX_BINS")) maxBins = strtoull(mb, nullptr, 10);  // tuning: bin entries for larger annotations
    const bool useEnt = fshift == 0 && (totalExtent >> 6) + 3ull * nChr <= maxBins && !getenv("MMANNOT_B200_NO_BINS");
    if (useEnt) fshift = 6;
    if (fshift == 0) {
      fshift = 5;
      while (fshift < 24 && (totalExtent >> fshift) > (1ull << 22)) ++fshift;
    }
    fshift = std::min<uint32_t>(std::max<uint32_t>(fshift, 5), 24);
    const uint32_t gshift = useEnt ? 0 : fshift - 5;
    std::vector<u32> fBase(nChr + 1, 0);
    std::vector<uint2> fInfo(nChr + 1);
    uint64_t fEntries = 0;
    for (uint32_t c = 0; c < nChr; ++c) {
      // (bin entries: one more bin, so that the last bin of a chromosome never holds a boundary)
      const uint64_t nb = (chrStart[c + 1] > chrStart[c]) ? (extent[c] >> fshift) + (useEnt ? 3 : 2) : 1;
      fBase[c] = (u32)fEntries;
      fInfo[c] = make_uint2((u32)fEntries, (u32)nb);
      fEntries += nb;
    }
    fBase[nChr] = (u32)fEntries;
    fInfo[nChr] = make_uint2((u32)fEntries, 1u);  // bin entries: the dummy bin of unknown chromosomes
    if (fEntries <= 0x7FFFFFFFull && nSeg < (1u << 30)) {
      CKS(segKey.ensure((size_t)nSeg * 8));
      CKS(ctx->fastSeg.ensure(((size_t)nSeg + 16) * 2 * sizeof(uint4)));
      CKS(ctx->fastTie.ensure(((size_t)nSeg + 16) * 2 * sizeof(u32)));
      u32 upMask = 0, downMask = 0;
      for (uint32_t q = 0; q < ctx->params.n_elements; ++q) {
        if (ctx->elemVic[q] == MMA_VICINITY_UP) upMask |= 1u << q;
        if (ctx->elemVic[q] == MMA_VICINITY_DOWN) downMask |= 1u << q;
      }
      if (useEnt) { CKS(ctx->fastEnt.ensure(((size_t)fEntries + 1) * sizeof(uint4))); CKS(ctx->fastRank.ensure(((size_t)fEntries + 1) * 4)); CKS(ctx->fastDict.ensure(ENT_DICT * sizeof(uint2))); }
      else CKS(ctx->fastBin.ensure((size_t)fEntries * sizeof(uint2)));
      CKS(ctx->fastChrInfo.ensure((size_t)(nChr + 1) * sizeof(uint2)));
      CKS(fChrBinBase.ensure((size_t)(nChr + 1) * 4));
      CKS(cudaMemsetAsync(ctx->fastSeg.p, 0, ((size_t)nSeg + 16) * 2 * sizeof(uint4), st));
      CKS(cudaMemcpyAsync(ctx->fastChrInfo.p, fInfo.data(), (size_t)(nChr + 1) * sizeof(uint2), cudaMemcpyHostToDevice, st));
      CKS(cudaMemcpyAsync(fChrBinBase.p, fBase.data(), (size_t)(nChr + 1) * 4, cudaMemcpyHostToDevice, st));
      {
        mma_ctx::Timed t(ctx, TC_INDEX);
        k_seg_scatter<<<gridFor(nKeys, 256), 256, 0, st>>>(kB.as<u64>(), flag.as<u32>(), pos.as<u32>(), (u32)nKeys, segKey.as<u64>());
        k_seg_eval<<<gridFor(nSeg, 128), 128, 0, st>>>(ctx->index, segKey.as<u64>(), nSeg, upMask, downMask, ctx->fastSeg.as<uint4>(), ctx->fastTie.as<u32>());
        if (useEnt) {
          CKS(dictHash.ensure(DICT_SLOTS * 12 + 4));
          CKS(cudaMemsetAsync(dictHash.p, 0, DICT_SLOTS * 12 + 4, st));
          BinBuild bb;
          bb.segKey = segKey.as<u64>(); bb.chrBinBase = fChrBinBase.as<u32>(); bb.seg = ctx->fastSeg.as<uint4>();
          bb.nSeg = nSeg; bb.nChr = nChr; bb.nEntries = (u32)fEntries; bb.nElements = ctx->params.n_elements;
          bb.hashKey = dictHash.as<u64>(); bb.hashId = reinterpret_cast<u32 *>(dictHash.as<u64>() + DICT_SLOTS);
          k_bin_entries<<<gridFor(fEntries + 1, 256), 256, 0, st>>>(bb, 0, ctx->fastEnt.as<uint4>(), ctx->fastRank.as<u32>());
          k_dict_number<<<1, 1024, 0, st>>>(bb.hashKey, bb.hashId, ctx->fastDict.as<uint2>(), bb.hashId + DICT_SLOTS);
          k_bin_entries<<<gridFor(fEntries + 1, 256), 256, 0, st>>>(bb, 1, ctx->fastEnt.as<uint4>(), ctx->fastRank.as<u32>());
          ctx->launches += 2;
          CKS(cudaMemcpyAsync(&nDictUsed, bb.hashId + DICT_SLOTS, 4, cudaMemcpyDeviceToHost, st));
        } else
          k_fast_bitmap<<<gridFor(fEntries, 256), 256, 0, st>>>(segKey.as<u64>(), nSeg, fChrBinBase.as<u32>(), nChr, fshift, gshift, (u32)fEntries,
                                                                ctx->fastBin.as<uint2>());
        ctx->launches += 3;
      }
      CKS(cudaStreamSynchronize(st));
      CKS(cudaGetLastError());
      ctx->fast.bm = useEnt ? nullptr : ctx->fastBin.as<uint2>();
      ctx->fast.ent = useEnt ? ctx->fastEnt.as<uint4>() : nullptr;
      ctx->fast.rank = useEnt ? ctx->fastRank.as<u32>() : nullptr;
      ctx->fast.dict = useEnt ? ctx->fastDict.as<uint2>() : nullptr;
      ctx->fast.nDict = useEnt ? std::max<u32>(2u, std::min<u32>(nDictUsed, ENT_DICT)) : 0u;
      ctx->fast.seg = ctx->fastSeg.as<uint4>();
      ctx->fast.tie = ctx->fastTie.as<u32>();
      ctx->fast.upMask = upMask;
      ctx->fast.downMask = downMask;
      ctx->fast.chrInfo = ctx->fastChrInfo.as<uint2>();
      ctx->fast.nChr = nChr;
      ctx->fast.shift = fshift;
      ctx->fast.gshift = gshift;
      ctx->fast.enabled = 1;
      ctx->nSegments = nSeg;
      fastBytes = (uint64_t)nSeg * (2 * sizeof(uint4) + 2 * sizeof(u32)) + fEntries * (useEnt ? sizeof(uint4) + 4 : sizeof(uint2)) + (uint64_t)nChr * sizeof(uint2);
    }
#undef CKS
    cleanupFast();
  }
#undef CKL
  cleanup();
  ctx->index.feat = ctx->feat.as<uint4>();
  ctx->index.chrInfo = ctx->chrInfo.as<uint2>();
  ctx->index.bins = ctx->bins.as<uint2>();
  ctx->index.spanIdx = ctx->spanIdx.as<u32>();
  ctx->index.nChr = nChr;
  ctx->index.shift = shift;
  ctx->indexBytes = (uint64_t)n * sizeof(uint4) + (uint64_t)nChr * sizeof(uint2) + entries * sizeof(uint2) + (uint64_t)total * 4 + fastBytes;
  ctx->haveIndex = true;
  return MMA_OK;
}

uint64_t mma_index_bytes(const mma_ctx *ctx) { return ctx ? ctx->indexBytes : 0; }
uint64_t mma_index_segments(const mma_ctx *ctx) { return ctx ? ctx->nSegments : 0; }
uint64_t mma_readback_bytes(const mma_ctx *ctx) { return ctx ? (uint64_t)std::min<u32>(ctx->tableCap, 8192u) * 16 + sizeof(TableDump) : 0; }
const char *mma_dominant_kernel(void) { return "k_batch_lean"; }
const char *mma_batch_kernel(const mma_ctx *ctx) {
  if (!ctx || !ctx->haveIndex) return "";
  if (ctx->wideMask || !ctx->fast.enabled || ctx->rules.strategy == MMA_STRATEGY_RANDOM || ctx->legacyBatch || ctx->fast.nChr > CHR_SMEM) return "k_batch";
  return ctx->fast.ent ? "k_batch_lean" : "k_batch_fast";
}

static int submitCommon(mma_ctx *ctx, uint32_t sample, const mma_hit_batch *b, bool onDevice) {
  int rc = checkSubmit(ctx, sample, b);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  if ((rc = initSample(ctx, s))) return rc;
  const uint64_t n = b->n;
  const uint64_t k = ctx->submitSeq;
  Staging &g = ctx->stage[k & 1];
  // buffers handed to the previous submit are free once its copy has landed
  if (k > 0) CK(cudaEventSynchronize(ctx->stage[(k - 1) & 1].copied));
  if (n == 0) return MMA_OK;
  if ((rc = ensureDeferred(ctx, s, n))) return rc;
  // this staging slot was last used by batch k-2
  if (k > 1) CK(cudaEventSynchronize(g.done));
  HitView h;
  h.n = (u32)n;
  if (onDevice) {
    h.start = b->start; h.end = b->end; h.meta = b->meta; h.nh = b->nh; h.key = (const u64 *)b->read_key;
    CK(cudaEventRecord(g.copied, ctx->sh));
  } else {
    const size_t cap = ctx->params.max_batch_hits;
    CK(g.start.ensure(cap * 4)); CK(g.end.ensure(cap * 4)); CK(g.meta.ensure(cap * 4)); CK(g.nh.ensure(cap * 4)); CK(g.key.ensure(cap * 8));
    CK(cudaMemcpyAsync(g.start.p, b->start, n * 4, cudaMemcpyHostToDevice, ctx->sh));
    CK(cudaMemcpyAsync(g.end.p, b->end, n * 4, cudaMemcpyHostToDevice, ctx->sh));
    CK(cudaMemcpyAsync(g.meta.p, b->meta, n * 4, cudaMemcpyHostToDevice, ctx->sh));
    CK(cudaMemcpyAsync(g.nh.p, b->nh, n * 4, cudaMemcpyHostToDevice, ctx->sh));
    CK(cudaMemcpyAsync(g.key.p, b->read_key, n * 8, cudaMemcpyHostToDevice, ctx->sh));
    CK(cudaEventRecord(g.copied, ctx->sh));
    CK(cudaStreamWaitEvent(ctx->sc, g.copied, 0));
    h.start = g.start.as<u32>(); h.end = g.end.as<u32>(); h.meta = g.meta.as<u32>(); h.nh = g.nh.as<u32>(); h.key = g.key.as<u64>();
  }
  h.vec = ((((uintptr_t)h.start | (uintptr_t)h.end | (uintptr_t)h.meta | (uintptr_t)h.nh | (uintptr_t)h.key) & 15u) == 0) ? 1u : 0u;
  if (!s.touched) s.deferAll = ctx->preferDefer && ctx->rules.strategy == MMA_STRATEGY_DEFAULT;  // (first batch of the sample)
  if ((rc = launchBatch(ctx, s, h))) return rc;
  CK(cudaEventRecord(g.done, ctx->sc));
  if ((rc = afterBatch(ctx, s, n))) return rc;
  ctx->submitSeq++;
  return MMA_OK;
}

int mma_pack_hits(const mma_hit_batch *w, uint32_t *packed, uint64_t *run_key, uint32_t *tile_run_base, uint32_t *esc_index,
                  uint32_t *esc_end, uint32_t *esc_nh, uint64_t esc_capacity, mma_packed_batch *out) {
  if (!w || !out || (w->n && (!w->start || !w->end || !w->meta || !w->nh || !w->read_key || !packed || !run_key || !tile_run_base)))
    return MMA_ERR_INVALID;
  uint64_t runs = 0, esc = 0;
  for (uint64_t i = 0; i < w->n; ++i) {
    if ((i % MMA_PACK_TILE) == 0) tile_run_base[i / MMA_PACK_TILE] = (uint32_t)runs;
    const uint32_t chrWide = w->meta[i] & MMA_HIT_CHR_MASK;
    uint32_t chr = MMA_PACKED_CHR_NONE;
    if (chrWide != MMA_HIT_CHR_NONE) {
      if (chrWide >= MMA_PACKED_CHR_NONE) return MMA_ERR_CAPACITY;
      chr = chrWide;
    }
    const uint32_t len = w->end[i] - w->start[i] + 1u;  // 0 for end = start - 1
    uint32_t lenField = len < 255u ? len : 255u, nhField = w->nh[i] < 255u ? w->nh[i] : 255u;
    if (lenField == 255u || nhField == 255u) {
      if (esc >= esc_capacity || !esc_index || !esc_end || !esc_nh) return MMA_ERR_CAPACITY;
      esc_index[esc] = (uint32_t)i; esc_end[esc] = w->end[i]; esc_nh[esc] = w->nh[i];
      ++esc;
    }
    uint32_t p = lenField | (nhField << 8) | (chr << 16) | (w->meta[i] & MMA_HIT_STRAND_BIT);
    // runs are defined on the normalised keys the kernels compare (the all-ones key is reserved, see normKey)
    const uint64_t kn = (w->read_key[i] == ~0ull) ? ~0ull - 1 : w->read_key[i];
    const uint64_t kp = (i == 0) ? 0 : ((w->read_key[i - 1] == ~0ull) ? ~0ull - 1 : w->read_key[i - 1]);
    if (i == 0 || kn != kp) { p |= MMA_PACKED_RUN_START; run_key[runs++] = kn; }
    packed[i] = p;
  }
  out->n = w->n; out->start = w->start; out->packed = packed; out->n_runs = runs; out->run_key = run_key;
  out->tile_run_base = tile_run_base; out->n_escapes = esc; out->esc_index = esc_index; out->esc_end = esc_end; out->esc_nh = esc_nh;
  return MMA_OK;
}

// Consistency of a batch in the compact format (no device work).  deep = 0: the pointers, the counts, tile_run_base (starts at 0,
// never falls, grows by at most a tile's hits per tile, stays within n_runs) and esc_index (strictly increasing, below n) --
// O(tiles + escapes), what mma_submit_hits_packed checks on every call.  deep != 0: also the run-start bits of every tile against
// tile_run_base / n_runs and every escaped hit against esc_index -- O(n), for callers that build the struct themselves.
int mma_check_packed(const mma_packed_batch *pb, int deep) {
  if (!pb) return MMA_ERR_INVALID;
  const uint64_t n = pb->n;
  if (n == 0) return MMA_OK;
  if (n > 0xFFFFFFF0ull) return MMA_ERR_INVALID;
  if (!pb->start || !pb->packed || !pb->run_key || !pb->tile_run_base || pb->n_runs == 0 || pb->n_runs > n) return MMA_ERR_INVALID;
  if (!(pb->packed[0] & MMA_PACKED_RUN_START)) return MMA_ERR_INVALID;
  if (pb->n_escapes > n || (pb->n_escapes && (!pb->esc_index || !pb->esc_end || !pb->esc_nh))) return MMA_ERR_INVALID;
  const uint64_t nTiles = (n + MMA_PACK_TILE - 1) / MMA_PACK_TILE;
  if (pb->tile_run_base[0] != 0) return MMA_ERR_INVALID;
  for (uint64_t t = 1; t < nTiles; ++t) {
    const uint64_t a = pb->tile_run_base[t - 1], b = pb->tile_run_base[t];
    if (b < a || b - a > MMA_PACK_TILE || b == 0 || b > pb->n_runs) return MMA_ERR_INVALID;
  }
  if (pb->n_runs - pb->tile_run_base[nTiles - 1] > n - (nTiles - 1) * MMA_PACK_TILE) return MMA_ERR_INVALID;
  for (uint64_t e = 0; e < pb->n_escapes; ++e)
    if (pb->esc_index[e] >= n || (e && pb->esc_index[e] <= pb->esc_index[e - 1])) return MMA_ERR_INVALID;
  if (deep) {
    uint64_t e = 0;
    for (uint64_t t = 0; t < nTiles; ++t) {
      const uint64_t lo = t * MMA_PACK_TILE, hi = std::min<uint64_t>(n, lo + MMA_PACK_TILE);
      uint64_t starts = 0;
      for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t pk = pb->packed[i];
        starts += (pk >> 30) & 1u;
        if ((pk & 255u) == 255u || ((pk >> 8) & 255u) == 255u) {  // an escaped hit must be listed
          while (e < pb->n_escapes && pb->esc_index[e] < i) ++e;
          if (e >= pb->n_escapes || pb->esc_index[e] != i) return MMA_ERR_INVALID;
        }
      }
      const uint64_t next = (t + 1 < nTiles) ? pb->tile_run_base[t + 1] : pb->n_runs;
      if (starts != next - pb->tile_run_base[t]) return MMA_ERR_INVALID;
    }
  }
  return MMA_OK;
}

int mma_submit_hits_packed(mma_ctx *ctx, uint32_t sample, const mma_packed_batch *pb) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!pb) return ctx->fail(MMA_ERR_INVALID, "null batch");
  if (!ctx->haveIndex) return ctx->fail(MMA_ERR_STATE, "mma_load_features must be called before hits are submitted");
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  if (pb->n > ctx->params.max_batch_hits) return ctx->fail(MMA_ERR_INVALID, "batch larger than max_batch_hits");
  const uint64_t n = pb->n;
  if (mma_check_packed(pb, 0) != MMA_OK) return ctx->fail(MMA_ERR_INVALID, "malformed packed batch");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  int rc = initSample(ctx, s);
  if (rc) return rc;
  const uint64_t k = ctx->submitSeq;
  Staging &g = ctx->stage[k & 1];
  if (k > 0) CK(cudaEventSynchronize(ctx->stage[(k - 1) & 1].copied));
  if (n == 0) return MMA_OK;
  if ((rc = ensureDeferred(ctx, s, n))) return rc;
  if (k > 1) CK(cudaEventSynchronize(g.done));
  const size_t cap = ctx->params.max_batch_hits;
  const size_t nTiles = (n + PACK_TILE - 1) / PACK_TILE;
  CK(g.start.ensure(cap * 4)); CK(g.end.ensure(cap * 4)); CK(g.meta.ensure(cap * 4)); CK(g.nh.ensure(cap * 4)); CK(g.key.ensure(cap * 8));
  CK(g.packed.ensure(cap * 4)); CK(g.runKey.ensure(cap * 8)); CK(g.tileBase.ensure(((cap + PACK_TILE - 1) / PACK_TILE) * 4));
  if (pb->n_escapes) { CK(g.escIndex.ensure(pb->n_escapes * 4)); CK(g.escEnd.ensure(pb->n_escapes * 4)); CK(g.escNh.ensure(pb->n_escapes * 4)); }
  CK(cudaMemcpyAsync(g.start.p, pb->start, n * 4, cudaMemcpyHostToDevice, ctx->sh));
  CK(cudaMemcpyAsync(g.packed.p, pb->packed, n * 4, cudaMemcpyHostToDevice, ctx->sh));
  CK(cudaMemcpyAsync(g.runKey.p, pb->run_key, pb->n_runs * 8, cudaMemcpyHostToDevice, ctx->sh));
  CK(cudaMemcpyAsync(g.tileBase.p, pb->tile_run_base, nTiles * 4, cudaMemcpyHostToDevice, ctx->sh));
  if (pb->n_escapes) {
    CK(cudaMemcpyAsync(g.escIndex.p, pb->esc_index, pb->n_escapes * 4, cudaMemcpyHostToDevice, ctx->sh));
    CK(cudaMemcpyAsync(g.escEnd.p, pb->esc_end, pb->n_escapes * 4, cudaMemcpyHostToDevice, ctx->sh));
    CK(cudaMemcpyAsync(g.escNh.p, pb->esc_nh, pb->n_escapes * 4, cudaMemcpyHostToDevice, ctx->sh));
  }
  CK(cudaEventRecord(g.copied, ctx->sh));
  CK(cudaStreamWaitEvent(ctx->sc, g.copied, 0));
  PackedView pv;
  pv.start = g.start.as<u32>(); pv.packed = g.packed.as<u32>(); pv.tileRunBase = g.tileBase.as<u32>();
  pv.escIndex = g.escIndex.as<u32>(); pv.escEnd = g.escEnd.as<u32>(); pv.escNh = g.escNh.as<u32>();
  pv.runKey = g.runKey.as<u64>(); pv.n = (u32)n; pv.nEsc = (u32)pb->n_escapes; pv.nRuns = (u32)pb->n_runs;
  {
    mma_ctx::Timed t(ctx, TC_CLOSE);
    k_expand_packed<<<(u32)nTiles, PACK_TILE / 4, 0, ctx->sc>>>(pv, g.end.as<u32>(), g.meta.as<u32>(), g.nh.as<u32>(), g.key.as<u64>());
    ctx->launches++;
  }
  HitView h;
  h.n = (u32)n;
  h.start = g.start.as<u32>(); h.end = g.end.as<u32>(); h.meta = g.meta.as<u32>(); h.nh = g.nh.as<u32>(); h.key = g.key.as<u64>();
  h.vec = 1u;
  if (!s.touched) s.deferAll = ctx->preferDefer && ctx->rules.strategy == MMA_STRATEGY_DEFAULT;  // (first batch of the sample)
  if ((rc = launchBatch(ctx, s, h))) return rc;
  CK(cudaEventRecord(g.done, ctx->sc));
  if ((rc = afterBatch(ctx, s, n))) return rc;
  ctx->submitSeq++;
  return MMA_OK;
}

int mma_submit_hits(mma_ctx *ctx, uint32_t sample, const mma_hit_batch *batch) { return submitCommon(ctx, sample, batch, false); }
int mma_submit_hits_device(mma_ctx *ctx, uint32_t sample, const mma_hit_batch *batch) { return submitCommon(ctx, sample, batch, true); }

int mma_annotate_hits(mma_ctx *ctx, const mma_hit_batch *b, uint64_t *out_masks) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!b || (b->n && (!b->start || !b->end || !b->meta || !out_masks))) return ctx->fail(MMA_ERR_INVALID, "null argument");
  if (!ctx->haveIndex) return ctx->fail(MMA_ERR_STATE, "mma_load_features must be called before hits are submitted");
  if (b->n == 0) return MMA_OK;
  if (b->n > 0xFFFFFFF0ull) return ctx->fail(MMA_ERR_INVALID, "too many hits in one call");
  CK(cudaSetDevice(ctx->device));
  const size_t n = b->n;
  DevBuf ds, de, dm, dout;
  auto cleanup = [&]() { ds.release(); de.release(); dm.release(); dout.release(); };
  cudaError_t e;
  if ((e = ds.ensure(n * 4)) != cudaSuccess || (e = de.ensure(n * 4)) != cudaSuccess || (e = dm.ensure(n * 4)) != cudaSuccess ||
      (e = dout.ensure(n * 8)) != cudaSuccess) { cleanup(); return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e)); }
  cudaStream_t st = ctx->sc;
  cudaMemcpyAsync(ds.p, b->start, n * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(de.p, b->end, n * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dm.p, b->meta, n * 4, cudaMemcpyHostToDevice, st);
  HitView h = HitView();
  h.start = ds.as<u32>(); h.end = de.as<u32>(); h.meta = dm.as<u32>(); h.n = (u32)n;
  const u32 grid = std::min<u32>(gridFor(n, 256), (u32)ctx->nSM * 8);
  const Rules &r = ctx->rules;
  const bool fast = ctx->fast.enabled && !ctx->wideMask;
#define LAUNCH_AO(M)                                                                                                        \
  do {                                                                                                                      \
    if (fast) k_annotate_only<M, true><<<grid, 256, 0, st>>>(ctx->index, ctx->fast, h, r, dout.as<u64>());                  \
    else k_annotate_only<M, false><<<grid, 256, 0, st>>>(ctx->index, ctx->fast, h, r, dout.as<u64>());                      \
  } while (0)
  if (r.mode == 0) LAUNCH_AO(0); else if (r.mode == 1) LAUNCH_AO(1); else LAUNCH_AO(2);
#undef LAUNCH_AO
  ctx->launches++;
  cudaMemcpyAsync(out_masks, dout.p, n * 8, cudaMemcpyDeviceToHost, st);
  e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e));
  return MMA_OK;
}

int mma_annotate_intervals(mma_ctx *ctx, const mma_hit_batch *b, uint64_t *out_masks, uint64_t *out_offsets, const uint32_t **out_ids) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!b || !out_offsets || !out_ids || (b->n && (!b->start || !b->end || !b->meta || !out_masks))) return ctx->fail(MMA_ERR_INVALID, "null argument");
  if (!ctx->haveIndex) return ctx->fail(MMA_ERR_STATE, "mma_load_features must be called before hits are submitted");
  *out_ids = nullptr;
  out_offsets[0] = 0;
  if (b->n == 0) return MMA_OK;
  if (b->n > 0x7FFFFFF0ull) return ctx->fail(MMA_ERR_INVALID, "too many hits in one call");
  CK(cudaSetDevice(ctx->device));
  const size_t n = b->n;
  DevBuf ds, de, dm, dmask, dcnt, doff, dtmp, dids;
  auto cleanup = [&]() { ds.release(); de.release(); dm.release(); dmask.release(); dcnt.release(); doff.release(); dtmp.release(); dids.release(); };
#define CKI(call)                                                                                  \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) { cleanup(); return ctx->fail(MMA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } \
  } while (0)
  CKI(ds.ensure(n * 4)); CKI(de.ensure(n * 4)); CKI(dm.ensure(n * 4)); CKI(dmask.ensure(n * 8)); CKI(dcnt.ensure(n * 4)); CKI(doff.ensure((n + 1) * 8));
  cudaStream_t st = ctx->sc;
  CKI(cudaMemcpyAsync(ds.p, b->start, n * 4, cudaMemcpyHostToDevice, st));
  CKI(cudaMemcpyAsync(de.p, b->end, n * 4, cudaMemcpyHostToDevice, st));
  CKI(cudaMemcpyAsync(dm.p, b->meta, n * 4, cudaMemcpyHostToDevice, st));
  HitView h = HitView();
  h.start = ds.as<u32>(); h.end = de.as<u32>(); h.meta = dm.as<u32>(); h.n = (u32)n;
  const u32 grid = std::min<u32>(gridFor(n, 256), (u32)ctx->nSM * 8);
  const Rules &r = ctx->rules;
  if (r.mode == 0) k_intervals_count<0><<<grid, 256, 0, st>>>(ctx->index, h, r, dmask.as<u64>(), dcnt.as<u32>());
  else if (r.mode == 1) k_intervals_count<1><<<grid, 256, 0, st>>>(ctx->index, h, r, dmask.as<u64>(), dcnt.as<u32>());
  else k_intervals_count<2><<<grid, 256, 0, st>>>(ctx->index, h, r, dmask.as<u64>(), dcnt.as<u32>());
  // exclusive scan of the per-hit counts (64-bit offsets) with the total in the extra last slot
  DevBuf dwide;
  auto cleanup2 = [&]() { dwide.release(); cleanup(); };
#define CKJ(call)                                                                                  \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) { cleanup2(); return ctx->fail(MMA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } \
  } while (0)
  CKJ(dwide.ensure((n + 1) * 8));
  CKJ(cudaMemsetAsync(dwide.p, 0, (n + 1) * 8, st));
  k_widen_counts<<<gridFor(n, 256), 256, 0, st>>>(dcnt.as<u32>(), dwide.as<u64>(), (u32)n);
  size_t tb = 0;
  CKJ(cub::DeviceScan::ExclusiveSum(nullptr, tb, dwide.as<u64>(), doff.as<u64>(), (int)(n + 1), st));
  CKJ(dtmp.ensure(tb ? tb : 1));
  CKJ(cub::DeviceScan::ExclusiveSum(dtmp.p, tb, dwide.as<u64>(), doff.as<u64>(), (int)(n + 1), st));
  ctx->launches += 2;
  CKJ(cudaMemcpyAsync(out_offsets, doff.p, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
  CKJ(cudaMemcpyAsync(out_masks, dmask.p, n * 8, cudaMemcpyDeviceToHost, st));
  CKJ(cudaStreamSynchronize(st));
  const uint64_t total = out_offsets[n];
  ctx->intervalIds.resize(total);
  if (total) {
    CKJ(dids.ensure(total * 4));
    if (r.mode == 0) k_intervals_fill<0><<<grid, 256, 0, st>>>(ctx->index, h, r, dmask.as<u64>(), doff.as<u64>(), dids.as<u32>());
    else if (r.mode == 1) k_intervals_fill<1><<<grid, 256, 0, st>>>(ctx->index, h, r, dmask.as<u64>(), doff.as<u64>(), dids.as<u32>());
    else k_intervals_fill<2><<<grid, 256, 0, st>>>(ctx->index, h, r, dmask.as<u64>(), doff.as<u64>(), dids.as<u32>());
    ctx->launches++;
    CKJ(cudaMemcpyAsync(ctx->intervalIds.data(), dids.p, total * 4, cudaMemcpyDeviceToHost, st));
    CKJ(cudaStreamSynchronize(st));
  }
  CKJ(cudaGetLastError());
#undef CKI
#undef CKJ
  cleanup2();
  *out_ids = ctx->intervalIds.data();
  return MMA_OK;
}

int mma_bam_begin(mma_ctx *ctx, uint32_t sample, const uint32_t *ref_to_chr, uint32_t n_ref, int strandedness) {
  if (!ctx) return MMA_ERR_INVALID;
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  if (n_ref && !ref_to_chr) return ctx->fail(MMA_ERR_INVALID, "null reference table");
  if (strandedness < 0 || strandedness > 2) return ctx->fail(MMA_ERR_INVALID, "strandedness must be 0 (U), 1 (F) or 2 (R)");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->bamRefToChr.ensure((size_t)std::max<u32>(n_ref, 1) * 4)); CK(ctx->bamRefFirst.ensure((size_t)std::max<u32>(n_ref, 1) * 8)); CK(ctx->bamFlags.ensure(8));
  if (n_ref) CK(cudaMemcpyAsync(ctx->bamRefToChr.p, ref_to_chr, (size_t)n_ref * 4, cudaMemcpyHostToDevice, ctx->sc));
  CK(cudaMemsetAsync(ctx->bamRefFirst.p, 0xFF, (size_t)std::max<u32>(n_ref, 1) * 8, ctx->sc));
  CK(cudaStreamSynchronize(ctx->sc));  // (ref_to_chr may be pageable: it must not change under the copy)
  ctx->bamNRef = n_ref;
  ctx->bamStrandedness = (u32)strandedness;
  ctx->bamOrdinal = 0;
  ctx->bamLastHits = 0;
  if (ctx->bamPend.active) {  // (a file given up half-way: its last chunk is dropped)
    CK(cudaStreamSynchronize(ctx->sc));
    for (int k = 0; k < 5; ++k) if (ctx->bamPend.ev[k]) ctx->eventPool.push_back(ctx->bamPend.ev[k]);
    ctx->bamPend = mma_ctx::BamPending();
  }
  for (int k = 0; k < 2; ++k)
    if (!ctx->bamCopied[k]) CK(cudaEventCreateWithFlags(&ctx->bamCopied[k], cudaEventDisableTiming));
  return MMA_OK;
}

static int bamGrow(mma_ctx *ctx, DevBuf &b, size_t need, size_t keep) {
  if (need <= b.bytes) return MMA_OK;
  DevBuf nb;
  CK(nb.ensure(std::max(need, b.bytes + b.bytes / 2)));
  if (keep) {
    CK(cudaStreamSynchronize(ctx->sh));
    CK(cudaMemcpy(nb.p, b.p, keep, cudaMemcpyDeviceToDevice));
  }
  b.release();
  b = nb;
  return MMA_OK;
}

int mma_bam_reserve(mma_ctx *ctx, uint64_t n_bytes) {
  if (!ctx) return MMA_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  // both staging slots: the one being filled, and the one the next chunk is staged into while this one is inflated
  int rc = bamGrow(ctx, ctx->bamComp[ctx->bamChunks & 1], (size_t)n_bytes + 16, (size_t)ctx->bamStaged);
  if (rc) return rc;
  return bamGrow(ctx, ctx->bamComp[(ctx->bamChunks & 1) ^ 1], (size_t)n_bytes + 16, 0);
}

int mma_bam_stage(mma_ctx *ctx, const void *host_bytes, uint64_t n_bytes, uint64_t offset) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!host_bytes && n_bytes) return ctx->fail(MMA_ERR_INVALID, "null argument");
  if (offset + n_bytes >= 0xFFFFFFF0ull) return ctx->fail(MMA_ERR_INVALID, "a BAM chunk must stay below 4 GB");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->bamStageEv) CK(cudaEventCreateWithFlags(&ctx->bamStageEv, cudaEventDisableTiming));
  else CK(cudaEventSynchronize(ctx->bamStageEv));  // the previous staged copy has left its host buffer
  DevBuf &dst = ctx->bamComp[ctx->bamChunks & 1];
  int rc = bamGrow(ctx, dst, (size_t)(offset + n_bytes) + 16, (size_t)ctx->bamStaged);
  if (rc) return rc;
  if (n_bytes) CK(cudaMemcpyAsync(dst.as<char>() + offset, host_bytes, (size_t)n_bytes, cudaMemcpyHostToDevice, ctx->sh));
  CK(cudaEventRecord(ctx->bamStageEv, ctx->sh));
  ctx->bamStaged = std::max<uint64_t>(ctx->bamStaged, offset + n_bytes);
  return MMA_OK;
}

// mma_submit_bam in two halves, so that the caller can read and stage the NEXT chunk while this one is inflated:
//   start   enqueues inflate -> record walk -> scan for the chunk (staged bytes or c->data) and returns at once
//   finish  waits for them, then parse -> batch kernels (the record count sizes the hit arrays and the launches)
// One chunk in flight at a time; mma_bam_stage calls between the two go to the other staging slot.
int mma_submit_bam_start(mma_ctx *ctx, uint32_t sample, const mma_bam_chunk *c) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!c) return ctx->fail(MMA_ERR_INVALID, "null argument");
  if (!ctx->haveIndex) return ctx->fail(MMA_ERR_STATE, "mma_load_features must be called before hits are submitted");
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  if (!ctx->bamFlags.p) return ctx->fail(MMA_ERR_STATE, "mma_bam_begin must be called first");
  if (ctx->bamPend.active) return ctx->fail(MMA_ERR_STATE, "a BAM chunk is already in flight (mma_submit_bam_finish)");
  ctx->bamPend = mma_ctx::BamPending();
  ctx->bamPend.sample = sample;
  if (c->n_members == 0) { ctx->bamPend.active = true; ctx->bamPend.empty = true; return MMA_OK; }
  if (!c->member_offset || !c->member_isize || c->n_bytes >= 0xFFFFFFF0ull || c->member_offset[c->n_members] != c->n_bytes)
    return ctx->fail(MMA_ERR_INVALID, "malformed BAM chunk");
  if (!c->data && c->n_bytes > ctx->bamStaged) return ctx->fail(MMA_ERR_INVALID, "fewer bytes staged (mma_bam_stage) than the chunk claims");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  int rc = initSample(ctx, s);
  if (rc) return rc;
  const u32 nM = c->n_members;
  std::vector<u32> outOff(nM + 1);
  uint64_t total = 0;
  for (u32 m = 0; m < nM; ++m) { outOff[m] = (u32)total; total += c->member_isize[m]; if (total >= 0xFFFFFFF0ull) return ctx->fail(MMA_ERR_INVALID, "BAM chunk inflates to 4 GB or more: submit fewer members"); }
  outOff[nM] = (u32)total;
  if (c->skip_first > c->member_isize[0]) return ctx->fail(MMA_ERR_INVALID, "skip_first beyond the first member");
  if (!ctx->bamPendHost) CK(cudaHostAlloc(&ctx->bamPendHost, 4 * sizeof(u32), cudaHostAllocDefault));
  if (!ctx->bamPendEv) CK(cudaEventCreateWithFlags(&ctx->bamPendEv, cudaEventDisableTiming));
  const int slot = (int)(ctx->bamChunks & 1);
  // (this slot's previous copy -- two chunks ago -- is long done: its kernels ran before the last chunk's, and that call synchronised)
  if (c->data) CK(ctx->bamComp[slot].ensure(((size_t)c->n_bytes + 15) & ~(size_t)15));
  CK(ctx->bamMemberOff.ensure((size_t)(nM + 1) * 4)); CK(ctx->bamOutOff.ensure((size_t)(nM + 1) * 4));
  CK(ctx->bamCount.ensure((size_t)nM * 4)); CK(ctx->bamHitOff.ensure((size_t)(nM + 1) * 4));
  if (c->data) CK(cudaMemcpyAsync(ctx->bamComp[slot].p, c->data, (size_t)c->n_bytes, cudaMemcpyHostToDevice, ctx->sh));
  CK(cudaEventRecord(ctx->bamCopied[slot], ctx->sh));
  ctx->bamStaged = 0;
  // the output buffer and the tables are shared by consecutive chunks: everything below is ordered on the compute stream (the
  // small tables first: a copy from pageable memory waits for the stream, which must not mean waiting for the staged bytes)
  CK(ctx->bamOut.ensure(((size_t)total + 64 + 15) & ~(size_t)15));
  CK(cudaMemcpyAsync(ctx->bamMemberOff.p, c->member_offset, (size_t)(nM + 1) * 4, cudaMemcpyHostToDevice, ctx->sc));
  CK(cudaMemcpyAsync(ctx->bamOutOff.p, outOff.data(), (size_t)(nM + 1) * 4, cudaMemcpyHostToDevice, ctx->sc));
  CK(cudaMemsetAsync(ctx->bamFlags.p, 0, 8, ctx->sc));  // (the flags word and the member counter of k_bam_inflate)
  CK(cudaStreamWaitEvent(ctx->sc, ctx->bamCopied[slot], 0));
  BamView &v = ctx->bamPend.v;
  v.comp = ctx->bamComp[slot].as<unsigned char>(); v.memberOff = ctx->bamMemberOff.as<u32>(); v.outOff = ctx->bamOutOff.as<u32>();
  v.out = ctx->bamOut.as<unsigned char>(); v.nMembers = nM; v.skipFirst = c->skip_first;
  v.refToChr = ctx->bamRefToChr.as<u32>(); v.nRef = ctx->bamNRef; v.strandedness = ctx->bamStrandedness;
  v.uniqueOnly = ctx->rules.strategy == MMA_STRATEGY_UNIQUE ? 1u : 0u;
  v.flags = ctx->bamFlags.as<u32>(); v.refFirst = ctx->bamRefFirst.as<unsigned long long>();
  cudaEvent_t *ev = ctx->bamPend.ev;
  if (ctx->timing) for (int k = 0; k < 5; ++k) ev[k] = ctx->getEvent();
  if (ev[0]) cudaEventRecord(ev[0], ctx->sc);
  {
    if (!ctx->bamInflateBlocks) {  // blocks the GPU holds at once
      int perSM = 0;
      CK(cudaFuncSetAttribute(k_bam_inflate, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));  // (its tables are its working set)
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_bam_inflate, BAM_INFLATE_THREADS, 0));
      ctx->bamInflateBlocks = (u32)std::max(1, perSM) * (u32)ctx->nSM;
    }
    const u32 warps = (nM + MMA_BAM_LANES - 1) / MMA_BAM_LANES;
    const u32 grid = std::min<u32>(gridFor((uint64_t)warps * 32, BAM_INFLATE_THREADS), ctx->bamInflateBlocks);
    k_bam_inflate<<<grid, BAM_INFLATE_THREADS, 0, ctx->sc>>>(v, ctx->bamFlags.as<u32>() + 1);
  }
  if (ev[1]) cudaEventRecord(ev[1], ctx->sc);
  k_bam_count<<<gridFor(nM, 128), 128, 0, ctx->sc>>>(v, ctx->bamCount.as<u32>());
  k_bam_scan<<<1, 1024, 0, ctx->sc>>>(ctx->bamCount.as<u32>(), nM, ctx->bamHitOff.as<u32>());
  if (ev[2]) cudaEventRecord(ev[2], ctx->sc);
  ctx->launches += 3;
  CK(cudaMemcpyAsync(&ctx->bamPendHost[0], ctx->bamFlags.p, 4, cudaMemcpyDeviceToHost, ctx->sc));
  CK(cudaMemcpyAsync(&ctx->bamPendHost[1], ctx->bamHitOff.as<u32>() + nM, 4, cudaMemcpyDeviceToHost, ctx->sc));
  CK(cudaEventRecord(ctx->bamPendEv, ctx->sc));
  ctx->bamChunks++;
  ctx->bamPend.active = true;
  return MMA_OK;
}

int mma_submit_bam_finish(mma_ctx *ctx, uint64_t *n_records, uint32_t *flags) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!n_records || !flags) return ctx->fail(MMA_ERR_INVALID, "null argument");
  if (!ctx->bamPend.active) return ctx->fail(MMA_ERR_STATE, "no BAM chunk in flight (mma_submit_bam_start)");
  *n_records = 0; *flags = 0;
  ctx->bamPend.active = false;
  if (ctx->bamPend.empty) return MMA_OK;
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[ctx->bamPend.sample];
  const BamView &v = ctx->bamPend.v;
  const u32 nM = v.nMembers;
  cudaEvent_t *ev = ctx->bamPend.ev;
  auto dropEvents = [&]() { for (int k = 0; k < 5; ++k) if (ev[k]) { ctx->eventPool.push_back(ev[k]); ev[k] = nullptr; } };
  CK(cudaEventSynchronize(ctx->bamPendEv));
  u32 hFlags = ctx->bamPendHost[0];
  const u32 nHits = ctx->bamPendHost[1];
  if (hFlags) { dropEvents(); *flags = hFlags; return MMA_OK; }
  if (nHits == 0) { dropEvents(); ctx->bamLastHits = 0; return MMA_OK; }
  const size_t cap = ((size_t)nHits + 127) & ~(size_t)127;
  CK(ctx->bamStart.ensure(cap * 4)); CK(ctx->bamEnd.ensure(cap * 4)); CK(ctx->bamMeta.ensure(cap * 4)); CK(ctx->bamNh.ensure(cap * 4)); CK(ctx->bamKey.ensure(cap * 8));
  HitOut o;
  o.start = ctx->bamStart.as<u32>(); o.end = ctx->bamEnd.as<u32>(); o.meta = ctx->bamMeta.as<u32>(); o.nh = ctx->bamNh.as<u32>(); o.key = ctx->bamKey.as<u64>();
  if (ev[3]) cudaEventRecord(ev[3], ctx->sc);
  k_bam_parse<<<gridFor(nM, 64), 64, 0, ctx->sc>>>(v, ctx->bamHitOff.as<u32>(), ctx->bamOrdinal, o);
  ctx->launches++;
  if (ev[4]) {
    cudaEventRecord(ev[4], ctx->sc);
    ctx->bamSpans.push_back({ev[0], ev[1], ev[2], ev[3], ev[4]});
    for (int k = 0; k < 5; ++k) ev[k] = nullptr;
  }
  // records with the flags only the parse can see (XA, odd CIGAR, odd aux)
  CK(cudaMemcpyAsync(&ctx->bamPendHost[0], ctx->bamFlags.p, 4, cudaMemcpyDeviceToHost, ctx->sc));
  CK(cudaStreamSynchronize(ctx->sc));
  hFlags = ctx->bamPendHost[0];
  if (hFlags) { *flags = hFlags; return MMA_OK; }
  ctx->bamOrdinal += nHits;
  ctx->bamLastHits = nHits;
  *n_records = nHits;
  // the hits are in HBM: annotate them like a device-resident batch, in pieces the per-thread counters can hold
  int rc;
  for (uint64_t a = 0; a < nHits;) {
    const uint64_t n = std::min<uint64_t>(nHits - a, 1ull << 27);
    if ((rc = ensureDeferred(ctx, s, n))) return rc;
    HitView h;
    h.n = (u32)n;
    h.start = o.start + a; h.end = o.end + a; h.meta = o.meta + a; h.nh = o.nh + a; h.key = o.key + a;
    h.vec = 1u;
    if (!s.touched) s.deferAll = ctx->preferDefer && ctx->rules.strategy == MMA_STRATEGY_DEFAULT;
    if ((rc = launchBatch(ctx, s, h))) return rc;
    if ((rc = afterBatch(ctx, s, n))) return rc;
    a += n;
  }
  return MMA_OK;
}

int mma_submit_bam(mma_ctx *ctx, uint32_t sample, const mma_bam_chunk *c, uint64_t *n_records, uint32_t *flags) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!c || !n_records || !flags) return ctx->fail(MMA_ERR_INVALID, "null argument");
  *n_records = 0; *flags = 0;
  const int rc = mma_submit_bam_start(ctx, sample, c);
  if (rc) return rc;
  return mma_submit_bam_finish(ctx, n_records, flags);
}

int mma_bam_ref_first(mma_ctx *ctx, uint64_t *out, uint32_t n_ref) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!out || n_ref != ctx->bamNRef) return ctx->fail(MMA_ERR_INVALID, "reference count differs from mma_bam_begin");
  if (n_ref == 0) return MMA_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->bamRefFirst.p, (size_t)n_ref * 8, cudaMemcpyDeviceToHost, ctx->sc));
  CK(cudaStreamSynchronize(ctx->sc));
  return MMA_OK;
}

int mma_bam_last_hits(mma_ctx *ctx, uint32_t *start, uint32_t *end, uint32_t *meta, uint32_t *nh, uint64_t *read_key) {
  if (!ctx) return MMA_ERR_INVALID;
  const size_t n = ctx->bamLastHits;
  if (n == 0) return MMA_OK;
  if (!start || !end || !meta || !nh || !read_key) return ctx->fail(MMA_ERR_INVALID, "null argument");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->sc));
  CK(cudaMemcpy(start, ctx->bamStart.p, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(end, ctx->bamEnd.p, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(meta, ctx->bamMeta.p, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(nh, ctx->bamNh.p, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(read_key, ctx->bamKey.p, n * 8, cudaMemcpyDeviceToHost));
  return MMA_OK;
}

int mma_sync(mma_ctx *ctx) {
  if (!ctx) return MMA_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->sh));
  CK(cudaStreamSynchronize(ctx->sc));
  return MMA_OK;
}

void *mma_stream(mma_ctx *ctx) { return ctx ? (void *)ctx->sc : nullptr; }

int mma_reset_sample(mma_ctx *ctx, uint32_t sample) {
  if (!ctx) return MMA_ERR_INVALID;
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  if (!s.ctl) return MMA_OK;
  const size_t tb = (size_t)ctx->tableCap * sizeof(u64);
  CK(cudaMemsetAsync(s.ctl, 0, sizeof(SampleCtl), ctx->sc));
  CK(cudaMemsetAsync(s.tableKeys.p, 0, tb, ctx->sc)); CK(cudaMemsetAsync(s.tableVals.p, 0, tb, ctx->sc));
  if (s.openCap && s.openMaybeUsed) {
    k_fill_u64<<<gridFor(s.openCap, 256), 256, 0, ctx->sc>>>(s.openKeys.as<u64>(), KEY_EMPTY, s.openCap);
    k_fill_u32<<<gridFor(s.openCap, 256), 256, 0, ctx->sc>>>(s.openSeq.as<u32>(), 0xFFFFFFFFu, s.openCap);
    ctx->launches += 2;
  }
  s.openMaybeUsed = false;  // (stream-ordered: the next batch kernels of this sample run after these fills)
  s.cumHits = 0; s.knownCount = 0; s.knownCum = 0; s.seq = 0; s.touched = false; s.deferAll = false;
  for (int i = 0; i < 4; ++i) s.ringUsed[i] = false;
  return MMA_OK;
}

static int finishDeferred(mma_ctx *ctx, Sample &s, u32 nSlow) {
  const Rules &r = ctx->rules;
  TableView table = tableView(s.tableKeys, s.tableVals, ctx->tableCap, s.ctl);
  SlowView slow = slowView(s);
  cudaStream_t st = ctx->sc;
  DevBuf &permA = ctx->defPermA, &permB = ctx->defPermB, &keyA = ctx->defKeyA, &keyB = ctx->defKeyB, &tmp = ctx->defTmp;
  auto cleanup = [&]() {};  // (the buffers stay with the context)
#define CKF(call)                                                                                  \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) { cleanup(); return ctx->fail(MMA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } \
  } while (0)
  CKF(permA.ensure((size_t)nSlow * 4)); CKF(permB.ensure((size_t)nSlow * 4));
  CKF(keyA.ensure((size_t)nSlow * 8)); CKF(keyB.ensure((size_t)nSlow * 8));
  size_t tmpBytes = 0;
  CKF(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keyA.as<u64>(), keyB.as<u64>(), permA.as<u32>(), permB.as<u32>(), (int)nSlow, 0, 64, st));
  CKF(tmp.ensure(tmpBytes));
  mma_ctx::Timed t(ctx, TC_FINISH);
  const u32 g = gridFor(nSlow, 256);
  if (r.strategy == MMA_STRATEGY_DEFAULT) {
    // ONE radix sort, by read key: the records of a name become adjacent (and are gathered into that order); each name's few
    // records are then taken in file order by selection (k_slow_default_byord).  Names with hundreds of records would make that
    // quadratic: then the list is sorted by (key, ordinal) below like for -y random.
    CKF(ctx->defOrd.ensure((size_t)nSlow * 8)); CKF(ctx->defMask.ensure((size_t)nSlow * 8)); CKF(ctx->defNh.ensure((size_t)nSlow * 4)); CKF(ctx->defMax.ensure(4));
    CKF(cudaMemsetAsync(ctx->defMax.p, 0, 4, st));
    k_iota<<<g, 256, 0, st>>>(permA.as<u32>(), nSlow);
    CKF(cudaMemcpyAsync(keyA.p, slow.key, (size_t)nSlow * 8, cudaMemcpyDeviceToDevice, st));
    CKF(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keyA.as<u64>(), keyB.as<u64>(), permA.as<u32>(), permB.as<u32>(), (int)nSlow, 0, 64, st));
    SlowView sorted;
    sorted.key = keyB.as<u64>(); sorted.ord = ctx->defOrd.as<u64>(); sorted.mask = ctx->defMask.as<u64>(); sorted.nh = ctx->defNh.as<u32>(); sorted.cap = nSlow;
    k_slow_gather<<<g, 256, 0, st>>>(permB.as<u32>(), nSlow, slow, sorted);
    k_slow_maxlen<<<g, 256, 0, st>>>(nSlow, sorted, ctx->defMax.as<u32>());
    u32 hMax = 0;
    CKF(cudaMemcpyAsync(&hMax, ctx->defMax.p, 4, cudaMemcpyDeviceToHost, st));
    CKF(cudaStreamSynchronize(st));
    ctx->launches += 3;
    if (hMax <= 512) {
      k_slow_default_byord<<<g, 256, 0, st>>>(nSlow, sorted, r, table, s.ctl);
      ctx->launches++;
      CKF(cudaStreamSynchronize(st));
      return MMA_OK;
    }
  }
  k_iota<<<g, 256, 0, st>>>(permA.as<u32>(), nSlow);
  CKF(cudaMemcpyAsync(keyA.p, slow.ord, (size_t)nSlow * 8, cudaMemcpyDeviceToDevice, st));
  CKF(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keyA.as<u64>(), keyB.as<u64>(), permA.as<u32>(), permB.as<u32>(), (int)nSlow, 0, 64, st));
  k_gather_keys<<<g, 256, 0, st>>>(permB.as<u32>(), nSlow, slow.key, keyA.as<u64>());
  CKF(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keyA.as<u64>(), keyB.as<u64>(), permB.as<u32>(), permA.as<u32>(), (int)nSlow, 0, 64, st));
  ctx->launches += 2;
  const u32 *perm = permA.as<u32>();
  if (r.strategy == MMA_STRATEGY_DEFAULT) {
    k_slow_default<<<g, 256, 0, st>>>(perm, nSlow, slow, r, table, s.ctl);
    ctx->launches++;
  } else {  // random
    DevBuf headOrd, headSorted, nHeadsDev, randDev;
    auto cleanup2 = [&]() { headOrd.release(); headSorted.release(); nHeadsDev.release(); randDev.release(); };
    cudaError_t e;
    if ((e = headOrd.ensure((size_t)nSlow * 8)) != cudaSuccess || (e = headSorted.ensure((size_t)nSlow * 8)) != cudaSuccess ||
        (e = nHeadsDev.ensure(4)) != cudaSuccess) { cleanup2(); cleanup(); return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e)); }
    // several input files: the names of this file join the map of the run (worst case: every deferred record another name)
    RndMap map;
    map.keys = nullptr; map.vals = nullptr; map.capMask = 0;
    if (ctx->params.n_samples > 1) {
      const uint64_t need = 2 * (ctx->rndNames + nSlow) + 1024;
      if (need > ctx->rndCap) {
        u32 cap = 1024;
        while (cap < need && cap < 0x80000000u) cap <<= 1;
        DevBuf nk, nv;
        if ((e = nk.ensure((size_t)cap * 8)) != cudaSuccess || (e = nv.ensure((size_t)cap * 8)) != cudaSuccess) { nk.release(); nv.release(); cleanup2(); cleanup(); return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e)); }
        k_fill_u64<<<gridFor(cap, 256), 256, 0, st>>>(nk.as<u64>(), KEY_EMPTY, cap);
        RndMap to;
        to.keys = nk.as<u64>(); to.vals = nv.as<u64>(); to.capMask = cap - 1;
        if (ctx->rndCap) {
          RndMap from;
          from.keys = ctx->rndKeys.as<u64>(); from.vals = ctx->rndVals.as<u64>(); from.capMask = ctx->rndCap - 1;
          k_rnd_rehash<<<gridFor(ctx->rndCap, 256), 256, 0, st>>>(from, to);
        }
        cudaStreamSynchronize(st);
        ctx->rndKeys.release(); ctx->rndVals.release();
        ctx->rndKeys = nk; ctx->rndVals = nv; ctx->rndCap = cap;
      }
      map.keys = ctx->rndKeys.as<u64>(); map.vals = ctx->rndVals.as<u64>(); map.capMask = ctx->rndCap - 1;
      ctx->rndNames += nSlow;
    }
    cudaMemsetAsync(nHeadsDev.p, 0, 4, st);
    k_slow_random_heads<<<g, 256, 0, st>>>(perm, nSlow, slow, map, headOrd.as<u64>(), nHeadsDev.as<u32>());
    ctx->launches++;
    u32 nHeads = 0;
    cudaMemcpyAsync(&nHeads, nHeadsDev.p, 4, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { cleanup2(); cleanup(); return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e)); }
    size_t tb2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb2, headOrd.as<u64>(), headSorted.as<u64>(), (int)nHeads, 0, 64, st);
    DevBuf tmp2;
    if ((e = tmp2.ensure(tb2 ? tb2 : 1)) != cudaSuccess) { cleanup2(); cleanup(); return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e)); }
    cub::DeviceRadixSort::SortKeys(tmp2.p, tb2, headOrd.as<u64>(), headSorted.as<u64>(), (int)nHeads, 0, 64, st);
    // the reference's rand() stream (glibc TYPE_3 additive feedback generator, mm:1711; never seeded => seed 1)
    // the draws of this sample follow those of the samples finished before it on this context: the reference never
    // reseeds between input files (mm:1711)
    const size_t skip = ctx->randDrawsUsed;
    std::vector<u32> stream((size_t)nHeads + skip + 344);
    {
      u32 seed = ctx->params.rand_seed ? ctx->params.rand_seed : 1u;
      std::vector<u32> &q = stream;
      q[0] = seed;
      for (int i = 1; i < 31; ++i) {
        int32_t hi = (int32_t)q[i - 1] / 127773, lo = (int32_t)q[i - 1] % 127773;
        int32_t w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        q[i] = (u32)w;
      }
      for (int i = 31; i < 34; ++i) q[i] = q[i - 31];
      for (size_t i = 34; i < q.size(); ++i) q[i] = q[i - 31] + q[i - 3];
      for (size_t i = 0; i < nHeads; ++i) q[i] = q[i + skip + 344] >> 1;
      ctx->randDrawsUsed += nHeads;
    }
    if ((e = randDev.ensure((size_t)std::max<u32>(nHeads, 1) * 4)) != cudaSuccess) { tmp2.release(); cleanup2(); cleanup(); return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e)); }
    cudaMemcpyAsync(randDev.p, stream.data(), (size_t)nHeads * 4, cudaMemcpyHostToDevice, st);
    k_slow_random_pick<<<g, 256, 0, st>>>(perm, nSlow, slow, headSorted.as<u64>(), nHeads, randDev.as<u32>(), r, table, map);
    ctx->launches++;
    e = cudaStreamSynchronize(st);
    tmp2.release();
    cleanup2();
    if (e != cudaSuccess) { cleanup(); return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e)); }
  }
  CKF(cudaStreamSynchronize(st));
#undef CKF
  cleanup();
  return MMA_OK;
}

// End-of-file flush of a sample (mm:1783-1792) and compaction of its table on the device: afterwards ctx->dumpBuf holds
// [TableDump | rows] and ctx->hostTable its head and (rows == true) rows.  One copy and one synchronisation in the usual
// case (no deferred records, a few thousand rows); `hc` receives the control block.
static int flushAndDump(mma_ctx *ctx, Sample &s, bool rows, SampleCtl &hc) {
  CK(cudaStreamSynchronize(ctx->sh));
  if (ctx->rules.strategy == MMA_STRATEGY_DEFAULT && s.slowCap) {
    k_flush_carry<<<1, 1, 0, ctx->sc>>>(s.ctl, slowView(s));
    ctx->launches++;
  }
  const size_t headBytes = sizeof(TableDump);
  const u32 firstRows = rows ? std::min<u32>(ctx->tableCap, 8192u) : 0u;
  if (!ctx->hostTable) CK(cudaHostAlloc(&ctx->hostTable, headBytes + (size_t)ctx->tableCap * 16, cudaHostAllocDefault));
  CK(ctx->dumpBuf.ensure(headBytes + (size_t)ctx->tableCap * 16));
  const TableDump *hHead = reinterpret_cast<const TableDump *>(ctx->hostTable);
  auto dump = [&]() -> int {
    TableDump *dHead = ctx->dumpBuf.as<TableDump>();
    ulonglong2 *dRows = reinterpret_cast<ulonglong2 *>(ctx->dumpBuf.as<char>() + headBytes);
    CK(cudaMemsetAsync(&dHead->nRows, 0, sizeof(u64), ctx->sc));
    k_table_compact<<<gridFor(ctx->tableCap, 256), 256, 0, ctx->sc>>>(tableView(s.tableKeys, s.tableVals, ctx->tableCap, s.ctl), s.ctl, dHead, dRows, ctx->tableCap);
    ctx->launches++;
    CK(cudaMemcpyAsync(ctx->hostTable, ctx->dumpBuf.p, headBytes + (size_t)firstRows * 16, cudaMemcpyDeviceToHost, ctx->sc));
    CK(cudaStreamSynchronize(ctx->sc));
    if (rows && hHead->nRows > firstRows) {
      const size_t off = headBytes + (size_t)firstRows * 16;
      CK(cudaMemcpyAsync(reinterpret_cast<char *>(ctx->hostTable) + off, ctx->dumpBuf.as<char>() + off, (size_t)(hHead->nRows - firstRows) * 16,
                         cudaMemcpyDeviceToHost, ctx->sc));
      CK(cudaStreamSynchronize(ctx->sc));
    }
    return MMA_OK;
  };
  int rc = dump();
  if (rc) return rc;
  hc = hHead->ctl;
  if (hc.overflow & 1u) return ctx->fail(MMA_ERR_CAPACITY, "a device table overflowed (combination table, deferred list or NH range under -y ratio)");
  if (hc.overflow & 2u) return ctx->fail(MMA_ERR_RETRY, "a shard of the exchange held deferred records: mma_restore_export, then the exchange with mma_export_table");
  s.openMaybeUsed = hc.openCount != 0;
  {  // what this sample says about the next one on this context (a sample of one batch never sees its own counters in time)
    const uint64_t hits = std::max<uint64_t>(hc.stats[ST_HITS], 1);
    if (ctx->forceGroups < 0 && ctx->rules.strategy == MMA_STRATEGY_DEFAULT && !s.deferAll) {
      if ((uint64_t)hc.walkCount * 32 > hits) ctx->useGroups = true;
      else if ((uint64_t)hc.walkCount * 256 < hits) ctx->useGroups = false;
    }
    if (ctx->forceDefer < 0 && ctx->rules.strategy == MMA_STRATEGY_DEFAULT && !s.deferAll && (uint64_t)hc.slowCount * 8 > hits) ctx->preferDefer = true;
  }
  if (hc.slowCount > 0) {
    if ((rc = finishDeferred(ctx, s, hc.slowCount))) return rc;
    // the deferred records are consumed: a later finish must not count them twice
    CK(cudaMemsetAsync(&s.ctl->slowCount, 0, sizeof(u32), ctx->sc));
    if ((rc = dump())) return rc;
    hc = hHead->ctl;
    if (hc.overflow) return ctx->fail(MMA_ERR_CAPACITY, "the combination table overflowed");
    s.knownCount = 0; s.knownCum = s.cumHits;
    // deferred for nothing: (nearly) every name's records were adjacent in the file -- the next sample runs the countdown again
    if (ctx->forceDefer < 0 && s.deferAll && (uint64_t)hc.deferContig * 10 >= (uint64_t)hc.deferSegs * 9) ctx->preferDefer = false;
  }
  return MMA_OK;
}

uint64_t mma_export_bytes(const mma_ctx *ctx) { return ctx ? sizeof(TableDump) + (uint64_t)ctx->tableCap * 16 : 0; }

int mma_export_table(mma_ctx *ctx, uint32_t sample, void *dev_dst) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!dev_dst) return ctx->fail(MMA_ERR_INVALID, "null destination");
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  int rc = initSample(ctx, s);
  if (rc) return rc;
  SampleCtl hc;
  if ((rc = flushAndDump(ctx, s, false, hc))) return rc;
  const u64 nRows = reinterpret_cast<const TableDump *>(ctx->hostTable)->nRows;
  CK(cudaMemcpyAsync(dev_dst, ctx->dumpBuf.p, sizeof(TableDump) + (size_t)nRows * 16, cudaMemcpyDeviceToDevice, ctx->sc));
  return MMA_OK;
}

// replaces the sample's table and counters by the sum over n_tables dumps ([TableDump | rows], `stride` bytes apart, at most
// `rowsCap` rows each looked at)
static int importTables(mma_ctx *ctx, Sample &s, const void *dev_src, uint32_t n_tables, size_t stride, u32 rowsCap) {
  const size_t tb = (size_t)ctx->tableCap * sizeof(u64);
  CK(cudaMemsetAsync(s.tableKeys.p, 0, tb, ctx->sc)); CK(cudaMemsetAsync(s.tableVals.p, 0, tb, ctx->sc));
  CK(cudaMemsetAsync(s.ctl->stats, 0, sizeof(u64) * ST_N, ctx->sc));
  CK(cudaMemsetAsync(&s.ctl->overflow, 0, sizeof(u32), ctx->sc));  // (the dumps bring their own flags; an import cut short is repeated)
  rowsCap = std::max<u32>(rowsCap, ST_N);  // (the first ST_N threads of each dump also add its counters)
  const u64 total = (u64)n_tables * rowsCap;
  k_table_import<<<gridFor(total, 256), 256, 0, ctx->sc>>>(tableView(s.tableKeys, s.tableVals, ctx->tableCap, s.ctl), s.ctl,
                                                          reinterpret_cast<const char *>(dev_src), n_tables, stride, rowsCap);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ctx->fail(MMA_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  return MMA_OK;
}

int mma_export_table_async(mma_ctx *ctx, uint32_t sample, void *dev_dst, uint64_t stride_bytes, uint64_t rows_cap) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!dev_dst) return ctx->fail(MMA_ERR_INVALID, "null destination");
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  if ((stride_bytes & 15) || rows_cap == 0 || rows_cap > ctx->tableCap || stride_bytes < sizeof(TableDump) + rows_cap * 16) return ctx->fail(MMA_ERR_INVALID, "bad stride / row capacity");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  int rc = initSample(ctx, s);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->sh));  // (host-side: the copies of the last batch have been handed to the copy engine)
  if (ctx->rules.strategy == MMA_STRATEGY_DEFAULT && s.slowCap) {
    k_flush_carry<<<1, 1, 0, ctx->sc>>>(s.ctl, slowView(s));
    ctx->launches++;
  }
  const size_t headBytes = sizeof(TableDump);
  CK(ctx->exportBuf.ensure(headBytes + (size_t)ctx->tableCap * 16));
  TableDump *dHead = ctx->exportBuf.as<TableDump>();
  ulonglong2 *dRows = reinterpret_cast<ulonglong2 *>(ctx->exportBuf.as<char>() + headBytes);
  CK(cudaMemsetAsync(&dHead->nRows, 0, sizeof(u64), ctx->sc));
  k_table_compact<<<gridFor(ctx->tableCap, 256), 256, 0, ctx->sc>>>(tableView(s.tableKeys, s.tableVals, ctx->tableCap, s.ctl), s.ctl, dHead, dRows, ctx->tableCap);
  ctx->launches++;
  CK(cudaMemcpyAsync(dev_dst, ctx->exportBuf.p, (size_t)stride_bytes, cudaMemcpyDeviceToDevice, ctx->sc));
  return MMA_OK;
}

int mma_restore_export(mma_ctx *ctx, uint32_t sample) {
  if (!ctx) return MMA_ERR_INVALID;
  if (sample >= ctx->samples.size() || !ctx->exportBuf.p) return ctx->fail(MMA_ERR_STATE, "nothing was exported");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  if (!s.ctl) return ctx->fail(MMA_ERR_STATE, "nothing was exported");
  int rc = importTables(ctx, s, ctx->exportBuf.p, 1, sizeof(TableDump) + (size_t)ctx->tableCap * 16, ctx->tableCap);
  if (rc) return rc;
  k_and_u32<<<1, 1, 0, ctx->sc>>>(&s.ctl->overflow, 1u);  // (the dump's own deferred records are still here: that is not an exchange gone wrong)
  ctx->launches++;
  return MMA_OK;
}

int mma_import_tables(mma_ctx *ctx, uint32_t sample, const void *dev_src, uint32_t n_tables) {
  return mma_import_tables_strided(ctx, sample, dev_src, n_tables, mma_export_bytes(ctx), ctx ? ctx->tableCap : 0);
}

int mma_import_tables_strided(mma_ctx *ctx, uint32_t sample, const void *dev_src, uint32_t n_tables, uint64_t stride_bytes, uint64_t rows_cap) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!dev_src || n_tables == 0) return ctx->fail(MMA_ERR_INVALID, "null source");
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  if ((stride_bytes & 15) || rows_cap == 0 || rows_cap > ctx->tableCap || stride_bytes < sizeof(TableDump) + rows_cap * 16)
    return ctx->fail(MMA_ERR_INVALID, "bad stride / row capacity");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  int rc = initSample(ctx, s);
  if (rc) return rc;
  return importTables(ctx, s, dev_src, n_tables, (size_t)stride_bytes, (u32)rows_cap);
}

uint64_t mma_export_head_bytes(void) { return sizeof(TableDump); }
uint64_t mma_export_rows(const mma_ctx *ctx) { return (ctx && ctx->hostTable) ? reinterpret_cast<const TableDump *>(ctx->hostTable)->nRows : 0; }

// ---- NCCL, opened at run time (a process that never merges over several GPUs never loads it; a process that already has a
//      libnccl.so.2 -- torch's -- gets that one)
namespace {
struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclCommInitAll) commInitAll = nullptr;
  decltype(&ncclCommDestroy) commDestroy = nullptr;
  decltype(&ncclAllGather) allGather = nullptr;
  decltype(&ncclGroupStart) groupStart = nullptr;
  decltype(&ncclGroupEnd) groupEnd = nullptr;
  decltype(&ncclGetErrorString) errorString = nullptr;
  std::string error;
  bool load() {
    if (handle) return true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (handle) break;
    }
    if (!handle) { error = std::string("cannot open libnccl.so.2: ") + dlerror(); return false; }
#define NCCL_SYM(field, sym)                                                          \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));                      \
  if (!field) { error = std::string("libnccl lacks ") + sym; handle = nullptr; return false; }
    NCCL_SYM(commInitAll, "ncclCommInitAll") NCCL_SYM(commDestroy, "ncclCommDestroy") NCCL_SYM(allGather, "ncclAllGather")
    NCCL_SYM(groupStart, "ncclGroupStart") NCCL_SYM(groupEnd, "ncclGroupEnd") NCCL_SYM(errorString, "ncclGetErrorString")
#undef NCCL_SYM
    return true;
  }
};
NcclApi g_nccl;
std::mutex g_ncclMutex;
std::map<std::vector<int>, std::vector<ncclComm_t>> g_comms;  // one clique per device list, kept for the life of the process
}  // namespace

int mma_allreduce(mma_ctx *const *ctxs, uint32_t n_ctx, uint32_t sample) {
  if (!ctxs || n_ctx == 0 || !ctxs[0]) return MMA_ERR_INVALID;
  mma_ctx *ctx = ctxs[0];  // (errors are reported on the first context)
  std::vector<int> devices;
  for (uint32_t i = 0; i < n_ctx; ++i) {
    if (!ctxs[i]) return ctx->fail(MMA_ERR_INVALID, "null context");
    if (sample >= ctxs[i]->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
    if (ctxs[i]->tableCap != ctx->tableCap || ctxs[i]->rules.strategy != ctx->rules.strategy) return ctx->fail(MMA_ERR_INVALID, "the contexts of a merge must share table size and strategy");
    if (std::find(devices.begin(), devices.end(), ctxs[i]->device) != devices.end()) return ctx->fail(MMA_ERR_INVALID, "one context per GPU");
    devices.push_back(ctxs[i]->device);
  }
  if (n_ctx == 1) return MMA_OK;
  std::lock_guard<std::mutex> lock(g_ncclMutex);
  if (!g_nccl.load()) return ctx->fail(MMA_ERR_STATE, g_nccl.error);
  std::vector<ncclComm_t> &comms = g_comms[devices];
  if (comms.empty()) {
    comms.resize(n_ctx);
    // (NCCL may print its version banner on stdout, where the command line writes the count table: send it to stderr)
    fflush(stdout);
    const int savedOut = dup(1);
    if (savedOut >= 0) dup2(2, 1);
    ncclResult_t r = g_nccl.commInitAll(comms.data(), (int)n_ctx, devices.data());
    fflush(stdout);
    if (savedOut >= 0) { dup2(savedOut, 1); close(savedOut); }
    if (r != ncclSuccess) { comms.clear(); g_comms.erase(devices); return ctx->fail(MMA_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.errorString(r)); }
  }
  // 1. end-of-file flush of every shard; its compacted table and counters into the context's dump buffer
  uint64_t maxRows = ST_N;
  for (uint32_t i = 0; i < n_ctx; ++i) {
    mma_ctx *c = ctxs[i];
    if (cudaSetDevice(c->device) != cudaSuccess) return ctx->fail(MMA_ERR_CUDA, "cudaSetDevice");
    Sample &s = c->samples[sample];
    int rc = initSample(c, s);
    SampleCtl hc;
    if (!rc) rc = flushAndDump(c, s, false, hc);
    if (rc) { if (c != ctx) ctx->fail(rc, c->error); return rc; }
    maxRows = std::max<uint64_t>(maxRows, reinterpret_cast<const TableDump *>(c->hostTable)->nRows);
  }
  // 2. ONE all-gather over NVLink of the live rows (padded to the longest shard), every GPU receiving every dump
  const size_t stride = (sizeof(TableDump) + (size_t)maxRows * 16 + 15) & ~(size_t)15;
  for (uint32_t i = 0; i < n_ctx; ++i) {
    mma_ctx *c = ctxs[i];
    cudaSetDevice(c->device);
    if (c->gatherBuf.ensure(stride * n_ctx) != cudaSuccess) return ctx->fail(MMA_ERR_CUDA, "cudaMalloc (gather buffer)");
  }
  ncclResult_t r = g_nccl.groupStart();
  for (uint32_t i = 0; i < n_ctx && r == ncclSuccess; ++i)
    r = g_nccl.allGather(ctxs[i]->dumpBuf.p, ctxs[i]->gatherBuf.p, stride, ncclUint8, comms[i], ctxs[i]->sc);
  ncclResult_t r2 = g_nccl.groupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) return ctx->fail(MMA_ERR_CUDA, std::string("ncclAllGather: ") + g_nccl.errorString(r));
  // 3. every GPU adds the dumps into its own (cleared) table: afterwards mma_finish_sample gives the merged result on any of them
  for (uint32_t i = 0; i < n_ctx; ++i) {
    mma_ctx *c = ctxs[i];
    cudaSetDevice(c->device);
    int rc = importTables(c, c->samples[sample], c->gatherBuf.p, n_ctx, stride, (u32)maxRows);
    if (rc) { if (c != ctx) ctx->fail(rc, c->error); return rc; }
  }
  return MMA_OK;
}

int mma_finish_sample(mma_ctx *ctx, uint32_t sample, mma_sample_result *out) {
  if (!ctx) return MMA_ERR_INVALID;
  if (!out) return ctx->fail(MMA_ERR_INVALID, "null result");
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  std::memset(out, 0, sizeof(*out));
  s.rowMask.clear(); s.rowNh.clear(); s.rowCount.clear();
  if (!s.ctl) return MMA_OK;  // nothing was ever submitted
  SampleCtl hc;
  int rc = flushAndDump(ctx, s, true, hc);
  if (rc) return rc;
  const size_t headBytes = sizeof(TableDump);
  const TableDump *hHead = reinterpret_cast<const TableDump *>(ctx->hostTable);
  const u64 nRows = hHead->nRows;
  const bool ratio = ctx->rules.strategy == MMA_STRATEGY_RATIO;
  const u64 lowMask = (1ull << NH_SHIFT) - 1;
  const u64 *hRows = reinterpret_cast<const u64 *>(reinterpret_cast<const char *>(ctx->hostTable) + headBytes);
  s.rowMask.reserve(nRows); s.rowNh.reserve(nRows); s.rowCount.reserve(nRows);
  for (u64 i = 0; i < nRows; ++i) {
    const u64 key = hRows[2 * i], val = hRows[2 * i + 1];
    s.rowMask.push_back(ratio ? (key & lowMask) : key);
    s.rowNh.push_back(ratio ? (uint32_t)(key >> NH_SHIFT) : 0u);
    s.rowCount.push_back(val);
  }
  out->stats.n_hits = hc.stats[ST_HITS];
  out->stats.n_reads = hc.stats[ST_READS];
  out->stats.n_unique = hc.stats[ST_UNIQUE];
  out->stats.n_ambiguous = hc.stats[ST_AMBIGUOUS];
  out->stats.n_multiple = hc.stats[ST_MULTIPLE];
  out->stats.n_unassigned = hc.stats[ST_UNASSIGNED];
  out->stats.n_rescued = hc.stats[ST_RESCUED];
  out->n_rows = s.rowMask.size();
  out->row_mask = s.rowMask.data();
  out->row_nh = s.rowNh.data();
  out->row_count = s.rowCount.data();
  return MMA_OK;
}

int mma_dense_counts(mma_ctx *ctx, uint32_t sample, const uint64_t *mask, const uint32_t *nh, uint64_t n, uint64_t *out_dev) {
  if (!ctx) return MMA_ERR_INVALID;
  if (sample >= ctx->samples.size()) return ctx->fail(MMA_ERR_INVALID, "sample index out of range");
  if (n == 0) return MMA_OK;
  if (!mask || !out_dev) return ctx->fail(MMA_ERR_INVALID, "null argument");
  CK(cudaSetDevice(ctx->device));
  Sample &s = ctx->samples[sample];
  int rc = initSample(ctx, s);
  if (rc) return rc;
  const bool ratio = ctx->rules.strategy == MMA_STRATEGY_RATIO;
  std::vector<u64> ck(n);
  for (uint64_t i = 0; i < n; ++i) ck[i] = mask[i] | ((ratio && nh) ? ((u64)nh[i] << NH_SHIFT) : 0ull);
  DevBuf d;
  CK(d.ensure(n * 8));
  CK(cudaMemcpyAsync(d.p, ck.data(), n * 8, cudaMemcpyHostToDevice, ctx->sc));
  k_dense_counts<<<gridFor(n, 256), 256, 0, ctx->sc>>>(tableView(s.tableKeys, s.tableVals, ctx->tableCap, s.ctl), d.as<u64>(), n, (u64 *)out_dev);
  ctx->launches++;
  cudaError_t e = cudaStreamSynchronize(ctx->sc);
  d.release();
  if (e != cudaSuccess) return ctx->fail(MMA_ERR_CUDA, cudaGetErrorString(e));
  return MMA_OK;
}

#ifdef MMA_DIAG
int mma_diag_get(unsigned long long *out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, mma::g_diag, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(mma::g_diag, z, sizeof(z)); }
  return 0;
}
#endif

int mma_timing_enable(mma_ctx *ctx, int on) {
  if (!ctx) return MMA_ERR_INVALID;
  ctx->timing = on != 0;
  return MMA_OK;
}
int mma_timing_reset(mma_ctx *ctx) {
  if (!ctx) return MMA_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->sc));
  ctx->collectTiming();
  for (int i = 0; i < TC_N; ++i) ctx->ms[i] = 0;
  for (int i = 0; i < 3; ++i) ctx->msBam[i] = 0;
  ctx->launches = 0;
  ctx->hitsSubmitted = 0;
  ctx->batches = 0;
  return MMA_OK;
}
int mma_timing_get(mma_ctx *ctx, mma_timing *out) {
  if (!ctx || !out) return MMA_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->sc));
  ctx->collectTiming();
  out->ms_index = ctx->ms[TC_INDEX];
  out->ms_batch = ctx->ms[TC_BATCH];
  out->ms_close = ctx->ms[TC_CLOSE];
  out->ms_finish = ctx->ms[TC_FINISH];
  out->launches = ctx->launches;
  out->hits = ctx->hitsSubmitted;
  out->batches = ctx->batches;
  out->fast_miss = 0;
  out->ms_bam_inflate = ctx->msBam[0]; out->ms_bam_index = ctx->msBam[1]; out->ms_bam_parse = ctx->msBam[2];
  for (Sample &sm : ctx->samples)
    if (sm.ctl) {
      u32 fm = 0;
      if (cudaMemcpy(&fm, &sm.ctl->fastMiss, sizeof(u32), cudaMemcpyDeviceToHost) == cudaSuccess) out->fast_miss += fm;
    }
  return MMA_OK;
}

}  // extern "C"
