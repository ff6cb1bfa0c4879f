// gpulin.cu -- host side of libgpulin.so: the C ABI of include/gpulin.h over the kernels of gpulin_kernels.cuh.
//
// Builds the device copy of the linear rows (row-binned SELL-32 + CSR, and the column -> row map), keeps the bound
// vectors resident, and drives the propagation rounds either from a CUDA graph with a device-side WHILE node
// (default: no host round trip per round) or from the host (GPULIN_LOOP=host; used for profiling single rounds).
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include "../../include/gpulin.h"
#include "gpulin_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

using namespace gpl;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(g_err, sizeof(g_err), fmt, ap);
   va_end(ap);
   return code;
}

#define CU(call)                                                                                                  \
   do                                                                                                             \
   {                                                                                                              \
      cudaError_t e_ = (call);                                                                                    \
      if( e_ != cudaSuccess )                                                                                     \
         return fail(GPULIN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
   } while( 0 )

#define OK(call)                 \
   do                            \
   {                             \
      int rc_ = (call);          \
      if( rc_ != GPULIN_OK )     \
         return rc_;             \
   } while( 0 )

// the read-only device arrays of the matrix, shared by a handle and its clones (gpulin_clone)
struct SharedMatrix
{
   void*  ptrs[64] = {nullptr};
   int    n = 0;
   int    refs = 1;
   size_t bytes = 0;
};

struct gpulin
{
   SharedMatrix* shared = nullptr;
   int         device = 0;
   int64_t     nrows = 0, ncols = 0, nnz = 0;
   int64_t     nstored = 0;      // nonzeros incl. SELL padding
   int         nsell = 0, nstream = 0, nlong = 0, ntiles = 0;
   int64_t     nstreamelems = 0;
   int         maxlen = 0;
   DevProblem  p{};
   int         nsellblocks = 0;
   int         nstreamblocks = 0;
   int         nlongblocks = 0;
   int         napplyblocks = 0;
   int         nexactblocks = 0;
   int         nfastblocks = 0;     // grid of fast_rows_kernel (0: no row of the thread-per-row class can take it)
   int         nsparseblocks = 0;   // grid of the persistent sparse-rounds kernel (0: disabled)
   void        (*exactkernel)(const DevProblem) = nullptr;
   unsigned char* d_redflags = nullptr; // gpulin_get_redundant_rows: one flag per row, device / pinned host (allocated on first use)
   unsigned char* h_redflags = nullptr;
   int         nsellunit = 0;       // SELL rows [0,nsellunit): all coefficients +1 / -1
   int         nsellbitsblocks = 0; // grid of sweep_sell_bits_kernel (0: the gather variant is used)
   int         sellbitsvariant = 0;
   size_t      freebytes = 0;       // size of the freebits array
   int         nsm = 148;
   std::vector<int> perm;        // permuted row -> caller's row
   // device allocations
   void*       d_all[48] = {nullptr};
   int         nalloc = 0;
   size_t      devbytes = 0;
   double*     d_tmplb = nullptr;   // staging for set/get_bounds
   double*     d_tmpub = nullptr;
   ChangeRec*  d_log = nullptr;
   int64_t     logcap = 0;
   Ctrl*       h_ctrl = nullptr;    // pinned mirror
   int*        h_params = nullptr;  // pinned {maxrounds, logcap}
   int*        d_updidx = nullptr;  // staging of gpulin_update_bounds
   double*     d_updlb = nullptr;
   double*     d_updub = nullptr;
   int64_t     updcap = 0;
   // ranged-row propagation (gpulin_set_rangedrow): host copy of the rows with two finite sides and >= 3 nonzeros
   std::vector<int> rr_row;             // permuted row id
   std::vector<long long> rr_ptr;
   std::vector<int> rr_col;
   std::vector<double> rr_val;
   std::vector<unsigned char> h_vartype;
   void*       d_rr[6] = {nullptr};     // device arrays of RangedRows (beg, row, cols, vals, idx, scratch)
   int         nrangedblocks = 0;
   double2*    d_ref = nullptr;     // reference bounds of gpulin_set_bounds_packed (allocated on first use)
   unsigned*   d_codes = nullptr;   // ... and the staged codes, 2 bits per column
   unsigned*   d_packlog = nullptr; // staging of gpulin_get_changes_packed (12 bytes per entry)
   int64_t     packlogcap = 0;
   unsigned*   d_clog = nullptr;    // staging of gpulin_get_changes_compact: count | words | explicit entries
   int64_t     clogcap = 0;
   unsigned*   h_clogcount = nullptr;  // pinned: number of explicit entries
   cudaStream_t stream = nullptr;
   bool        ownstream = true;
   cudaStream_t aux[2] = {nullptr, nullptr};     // the medium / long sweeps run beside the short sweep
   cudaEvent_t evfork = nullptr, evjoin[2] = {nullptr, nullptr};
   cudaEvent_t evfork2 = nullptr, evjoin2 = nullptr;   // exact_rows_kernel || fast_rows_kernel
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   cudaEvent_t evprof[4] = {nullptr, nullptr, nullptr, nullptr};
   cudaGraph_t graph = nullptr;
   cudaGraphExec_t gexec = nullptr;
   cudaGraphConditionalHandle handle = 0;
   bool        hostloop = false;
   int         npeers = 1;          // > 1 after gpulin_peer_connect / gpulin_group_connect: dense rounds are shared with the peers
   int         peerrank = 0;
   PeerTable*  d_peers = nullptr;
   unsigned char* d_inbox = nullptr; // this rank's inbox (exported to the peers)
   int         inboxranks = 0;      // ... sized for this many ranks
   void*       peerptr[MAX_PEERS] = {};   // opened IPC pointers of the other ranks' inboxes (multi-process connection)
   int         npushblocks = 0;
   bool        havebounds = false;
   bool        pending = false;     // gpulin_propagate_async was called, gpulin_propagate_wait not yet
   bool        lightfetch = false;  // (unused: every call fetches only the head of the control block now)
   bool        histfetched = false; // the pinned mirror holds the per-round history of the last call
   std::vector<gpulin*> workers;    // clones that gpulin_probe_batch keeps for this (base) handle
   void*       d_proberes = nullptr; // verdicts of the probes of a batch (ProbeResult[proberescap])
   int*        d_probevar = nullptr; // the probes of a batch on the device
   double*     d_probelb = nullptr;
   double*     d_probeub = nullptr;
   int64_t     proberescap = 0;
   ChangeRec*  d_probelog = nullptr;   // change logs of the probes of a batch (gpulin_probe_batch_changes) + the cursor behind them
   unsigned long long* d_probecursor = nullptr;
   int64_t     probelogcap = 0;
   uint64_t    version = 1;         // counts the calls that can change this handle's bounds (workers compare it)
   uint64_t    syncedversion = 0;   // probing worker: version of the base handle it last took the bounds from
   int64_t     smallcols = -1;      // columns updated since the last clean fixpoint (< 0: the marks are not all on the list)
   bool        smallcall = false;   // the pending call was started by probe_kernel
   bool        smallcalls = true;   // GPULIN_SMALL=0 disables that path
   bool        lastsmall = false;   // the last call was started by probe_kernel ...
   bool        lastresumed = false; // ... and handed over to the general loop after smallrounds rounds
   int         smallrounds = 0;
   int         lastmaxrounds = 0;
   int         lastvar = -1;        // probing worker: the column of its last probe
   bool        needreset = true;    // probing worker: its state is not "node + change log"
   // results of the last propagate call
   gpulin_result last{};
   int         lastrounds = 0;
};

template <typename T>
static int devAlloc(gpulin* h, T** out, size_t count, bool shared = false)
{
   void* ptr = nullptr;
   size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
   cudaError_t e = cudaMalloc(&ptr, bytes);
   if( e != cudaSuccess )
      return fail(GPULIN_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
   if( shared )
   {
      h->shared->ptrs[h->shared->n++] = ptr;
      h->shared->bytes += bytes;
   }
   else
      h->d_all[h->nalloc++] = ptr;
   h->devbytes += bytes;
   *out = (T*)ptr;
   return GPULIN_OK;
}

extern "C" void gpulin_default_numerics(gpulin_numerics* num)
{
   num->infinity = 1e20;
   num->epsilon = 1e-9;
   num->sumepsilon = 1e-6;
   num->feastol = 1e-6;
   num->boundstreps = 0.05;
   num->hugeval = 1e15;
   num->maxeasyactivitydelta = 1e6;
}

extern "C" const char* gpulin_last_error(void)
{
   return g_err;
}

extern "C" int gpulin_device_count(void)
{
   int n = 0;
   cudaError_t e = cudaGetDeviceCount(&n);
   if( e != cudaSuccess )
      return fail(GPULIN_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
   return n;
}

static void destroyGraph(gpulin* h)
{
   if( h->gexec != nullptr )
      cudaGraphExecDestroy(h->gexec);
   if( h->graph != nullptr )
      cudaGraphDestroy(h->graph);
   h->gexec = nullptr;
   h->graph = nullptr;
}

// dynamic shared memory of sweep_sell_bits_kernel: bit table, mbarrier
static size_t sellBitsSmem(int nfreewords)
{
   return (((size_t)nfreewords * 4 + 127) & ~(size_t)127) + SB_AUX_BYTES;
}

// the thread-per-row sweep: the gather variant (small instances), and the two instances of the bit-table sweep
// (1024 threads, two nonzeros per thread and chunk, four in the unit slices; the table covers a prefix of / all columns)
typedef void (*SweepKernel)(const DevProblem);
static const SweepKernel g_sellKernel = sweep_sell_kernel<2, 4>;
static const SweepKernel g_sellBitsKernels[2] = {
   sweep_sell_bits_kernel<1024, 2, false, 4>,
   sweep_sell_bits_kernel<1024, 2, true, 4>,
};
constexpr int SELLBITS_THREADS = 1024;

constexpr int64_t SMALLCALL_MAXCOLS = 256;      // gpulin_propagate after at most this many updated columns starts in one block

// one propagation round on h->stream
// the changed-column bits of the exact kernel become the change list; with peers this is the one exchange of a dense
// round: the listed columns go out to every peer, theirs are merged into the local keys
static int launchCollect(gpulin* h)
{
   if( h->npeers > 1 )
   {
      collect_kernel<true><<<h->npushblocks, 256, COLLECT_SMEM_PEERS, h->stream>>>(h->p);
      peer_merge_kernel<<<h->nsm * 8, 256, 0, h->stream>>>(h->p);
   }
   else
      collect_kernel<false><<<h->npushblocks, 256, 0, h->stream>>>(h->p);
   CU(cudaGetLastError());
   return GPULIN_OK;
}

// the exact rules for the rows the filter sweeps handed over; the rows that came with their activities run beside the
// others (fast_rows_kernel on a side stream: a parallel branch of the graph)
static int launchExact(gpulin* h, bool firstround)
{
   const bool fast = firstround && h->nfastblocks > 0;
   h->p.fastround = fast ? 1 : 0;      // (the kernels take the problem by value: this launch's copy says who takes flist)
   if( fast )
   {
      CU(cudaEventRecord(h->evfork2, h->stream));
      CU(cudaStreamWaitEvent(h->aux[0], h->evfork2, 0));
      fast_rows_kernel<<<h->nfastblocks, FAST_THREADS, 0, h->aux[0]>>>(h->p);
      CU(cudaEventRecord(h->evjoin2, h->aux[0]));
   }
   h->exactkernel<<<h->nexactblocks, EXACT_THREADS, 0, h->stream>>>(h->p);
   if( fast )
      CU(cudaStreamWaitEvent(h->stream, h->evjoin2, 0));
   h->p.fastround = 0;
   return GPULIN_OK;
}

template <int MODE, bool GRAPH>
static int launchRoundKernels(gpulin* h, bool sweep, bool apply, bool collect = true, bool firstround = false)
{
   if( sweep )
   {
      // ranged rows first: the gcd rule looks at the marks the sweeps are about to lower
      if( h->p.rr.n > 0 && MODE == APPLY_LIST )
         rangedrow_kernel<<<h->nrangedblocks, RANGED_THREADS, 0, h->stream>>>(h->p);
      // the three bins are independent: the smaller ones run on side streams beside the largest
      const int nkinds = (h->nsellblocks > 0) + (h->nstreamblocks > 0) + (h->nlongblocks > 0);
      int side = 0;
      if( nkinds > 1 )
         CU(cudaEventRecord(h->evfork, h->stream));
      if( h->nlongblocks > 0 )
      {
         cudaStream_t st = nkinds > 1 ? h->aux[side] : h->stream;
         if( nkinds > 1 )
            CU(cudaStreamWaitEvent(st, h->evfork, 0));
         sweep_long_kernel<<<h->nlongblocks, LONG_THREADS, 0, st>>>(h->p);
         if( nkinds > 1 )
         {
            CU(cudaEventRecord(h->evjoin[side], st));
            ++side;
         }
      }
      if( h->nstreamblocks > 0 && h->nsellblocks > 0 )
      {
         cudaStream_t st = h->aux[side];
         CU(cudaStreamWaitEvent(st, h->evfork, 0));
         sweep_stream_kernel<<<h->nstreamblocks, SWEEP_THREADS, 0, st>>>(h->p);
         CU(cudaEventRecord(h->evjoin[side], st));
         ++side;
      }
      else if( h->nstreamblocks > 0 )
         sweep_stream_kernel<<<h->nstreamblocks, SWEEP_THREADS, 0, h->stream>>>(h->p);
      if( h->nsellbitsblocks > 0 )
         g_sellBitsKernels[h->sellbitsvariant]<<<h->nsellbitsblocks, SELLBITS_THREADS, sellBitsSmem(h->p.nfreewords), h->stream>>>(h->p);
      else if( h->nsellblocks > 0 )
         g_sellKernel<<<h->nsellblocks, SELL_THREADS, 0, h->stream>>>(h->p);
      for( int i = 0; i < side; ++i )
         CU(cudaStreamWaitEvent(h->stream, h->evjoin[i], 0));
      if( h->nexactblocks > 0 )
         OK(launchExact(h, firstround));
      if( collect )
         OK(launchCollect(h));
   }
   if( apply )
   {
      apply_kernel<MODE, GRAPH><<<h->napplyblocks, APPLY_THREADS, 0, h->stream>>>(h->p, h->handle);
      if( MODE == APPLY_LIST && sweep && h->nsparseblocks > 0 )
      {
         // rounds with few marked rows are run by one persistent cooperative kernel (returns at once otherwise)
         void* args[2] = {(void*)&h->p, (void*)&h->handle};
         CU(cudaLaunchCooperativeKernel((const void*)sparse_rounds_kernel<GRAPH>, dim3(h->nsparseblocks), dim3(SPARSE_THREADS),
            args, 0, h->stream));
      }
   }
   CU(cudaGetLastError());
   return GPULIN_OK;
}

// graph:  begin_kernel -> WHILE(cont) { sweep kernels; apply kernel (sets cont) }
static int buildGraph(gpulin* h)
{
   destroyGraph(h);
   CU(cudaGraphCreate(&h->graph, 0));
   CU(cudaGraphConditionalHandleCreate(&h->handle, h->graph, 1, cudaGraphCondAssignDefault));

   cudaGraphNode_t beginNode;
   {
      cudaKernelNodeParams kp;
      memset(&kp, 0, sizeof(kp));
      void* args[1] = {(void*)&h->p.ctrl};
      kp.func = (void*)begin_kernel;
      kp.gridDim = dim3(1);
      kp.blockDim = dim3(1);
      kp.kernelParams = args;
      CU(cudaGraphAddKernelNode(&beginNode, h->graph, nullptr, 0, &kp));
   }
   // the first round in front of the loop: the one round that launches fast_rows_kernel (see DevProblem::fastround); its
   // apply step sets the condition of the loop like every other
   std::vector<cudaGraphNode_t> afterFirst;
   {
      CU(cudaStreamBeginCaptureToGraph(h->stream, h->graph, &beginNode, nullptr, 1, cudaStreamCaptureModeThreadLocal));
      const int frc = launchRoundKernels<APPLY_LIST, true>(h, true, true, true, true);
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      const cudaGraphNode_t* deps = nullptr;
      size_t ndeps = 0;
      cudaError_t ce = cudaStreamGetCaptureInfo(h->stream, &st, nullptr, nullptr, &deps, &ndeps);
      if( ce == cudaSuccess )
         afterFirst.assign(deps, deps + ndeps);
      cudaGraph_t captured1 = nullptr;
      CU(cudaStreamEndCapture(h->stream, &captured1));
      OK(frc);
      CU(ce);
      if( afterFirst.empty() )
         return fail(GPULIN_ERR_CUDA, "the capture of the first round left no node to continue from");
   }
   cudaGraphNodeParams cp = {};
   cp.type = cudaGraphNodeTypeConditional;
   cp.conditional.handle = h->handle;
   cp.conditional.type = cudaGraphCondTypeWhile;
   cp.conditional.size = 1;
   cudaGraphNode_t whileNode;
   CU(cudaGraphAddNode(&whileNode, h->graph, afterFirst.data(), afterFirst.size(), &cp));
   cudaGraph_t body = cp.conditional.phGraph_out[0];

   CU(cudaStreamBeginCaptureToGraph(h->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
   const int lrc = launchRoundKernels<APPLY_LIST, true>(h, true, true);
   cudaGraph_t captured = nullptr;
   CU(cudaStreamEndCapture(h->stream, &captured));
   OK(lrc);
   CU(cudaGraphInstantiate(&h->gexec, h->graph, 0));
   return GPULIN_OK;
}

// the share of rank `rank` of `nranks` in a dense round: every nranks-th SELL slice (the kernels interleave), an equal
// number of tiles of the stream, every nranks-th block-per-row row
static void setShare(gpulin* h, int rank, int nranks)
{
   DevProblem& p = h->p;
   p.st0 = (int)((long long)h->ntiles * rank / nranks);
   p.st1 = (int)((long long)h->ntiles * (rank + 1) / nranks);
   p.nranks = nranks;
   p.rank = rank;
}

extern "C" int gpulin_create(int device, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* rowptr,
   const int32_t* colidx, const double* vals, const double* lhs, const double* rhs, const uint8_t* vartype,
   const gpulin_numerics* num, gpulin_t** out)
{
   if( out == nullptr )
      return fail(GPULIN_ERR_ARG, "out is NULL");
   *out = nullptr;
   if( nrows < 0 || ncols < 0 || nnz < 0 || (nrows > 0 && rowptr == nullptr) || (nnz > 0 && (colidx == nullptr || vals == nullptr))
      || (nrows > 0 && (lhs == nullptr || rhs == nullptr)) || (ncols > 0 && vartype == nullptr) )
      return fail(GPULIN_ERR_ARG, "invalid problem arrays");
   if( nrows >= (1LL << 30) || ncols >= (1LL << 30) )
      return fail(GPULIN_ERR_ARG, "more than 2^30 rows or columns are not supported");
   if( nrows > 0 && (rowptr[0] != 0 || rowptr[nrows] != nnz) )
      return fail(GPULIN_ERR_ARG, "rowptr does not span [0,nnz]");
   for( int64_t r = 0; r < nrows; ++r )
   {
      if( rowptr[r + 1] < rowptr[r] )
         return fail(GPULIN_ERR_ARG, "rowptr is not monotone at row %lld", (long long)r);
      if( rowptr[r + 1] - rowptr[r] >= (1LL << 29) )
         return fail(GPULIN_ERR_ARG, "row %lld is too long", (long long)r);
   }
   for( int64_t k = 0; k < nnz; ++k )
   {
      if( colidx[k] < 0 || colidx[k] >= ncols )
         return fail(GPULIN_ERR_ARG, "column index %d out of range at position %lld", colidx[k], (long long)k);
      if( vals[k] == 0.0 || vals[k] != vals[k] )
         return fail(GPULIN_ERR_ARG, "zero or NaN coefficient at position %lld (cons_linear.c:5422)", (long long)k);
   }

   {
      int ndev = 0;
      cudaError_t e = cudaGetDeviceCount(&ndev);
      if( e != cudaSuccess || ndev <= 0 )
         return fail(GPULIN_ERR_CUDA, "no CUDA device available (%s); libgpulin has no CPU fallback",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
      if( device < 0 || device >= ndev )
         return fail(GPULIN_ERR_ARG, "device %d out of range [0,%d)", device, ndev);
   }
   CU(cudaSetDevice(device));
   gpulin_numerics defnum;
   gpulin_default_numerics(&defnum);
   if( num == nullptr )
      num = &defnum;

   gpulin* h = new gpulin();
   h->shared = new SharedMatrix();
   h->device = device;
   h->nrows = nrows;
   h->ncols = ncols;
   h->nnz = nnz;
   {
      int nsm = 0;
      if( cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && nsm > 0 )
         h->nsm = nsm;
   }
   h->smallcalls = !(getenv("GPULIN_SMALL") != nullptr && atoi(getenv("GPULIN_SMALL")) == 0);
   const char* loopenv = getenv("GPULIN_LOOP");
   h->hostloop = (loopenv != nullptr && strcmp(loopenv, "host") == 0);

   // ---- bin the rows ------------------------------------------------------------------------------------------------
   //   1..32 nonzeros       SELL-32 slices, sorted by length (stable: neighbours stay neighbours)   thread-per-row sweep
   //   33..STREAM_MAXLEN    one CSR stream in the caller's order, cut into tiles of 256 nonzeros     tile sweep
   //   longer, or empty     CSR, longest first                                                      block-per-row sweep
   std::vector<int> len((size_t)nrows);
   for( int64_t r = 0; r < nrows; ++r )
   {
      len[(size_t)r] = (int)(rowptr[r + 1] - rowptr[r]);
      h->maxlen = std::max(h->maxlen, len[(size_t)r]);
   }
   std::vector<int>& perm = h->perm;
   perm.clear();
   perm.reserve((size_t)nrows);
   std::vector<int> streamrows;
   std::vector<int> longrows;
   for( int64_t r = 0; r < nrows; ++r )
   {
      const int l = len[(size_t)r];
      if( l >= 1 && l <= SHORT_MAXLEN )
         perm.push_back((int)r);
      else if( l > SHORT_MAXLEN && l <= STREAM_MAXLEN )
         streamrows.push_back((int)r);
      else
         longrows.push_back((int)r);
   }
   // the SELL rows whose coefficients are all +1 or -1 come first, in whole slices (the filter sweep does not read their
   // values, see sweep_sell_bits_kernel); what does not fill a slice stays with the others
   {
      std::vector<int> unitrows;
      std::vector<int> otherrows;
      for( int r : perm )
      {
         bool unit = true;
         for( int64_t k = rowptr[r]; k < rowptr[r + 1] && unit; ++k )
            unit = std::fabs(vals[k]) == 1.0;
         (unit ? unitrows : otherrows).push_back(r);
      }
      auto bylen = [&](int a, int b) { return len[(size_t)a] < len[(size_t)b]; };
      std::stable_sort(unitrows.begin(), unitrows.end(), bylen);
      h->nsellunit = (int)(unitrows.size() / 32 * 32);
      otherrows.insert(otherrows.end(), unitrows.begin() + h->nsellunit, unitrows.end());
      std::stable_sort(otherrows.begin(), otherrows.end(), bylen);
      perm.assign(unitrows.begin(), unitrows.begin() + h->nsellunit);
      perm.insert(perm.end(), otherrows.begin(), otherrows.end());
   }
   std::stable_sort(longrows.begin(), longrows.end(), [&](int a, int b) { return len[(size_t)a] > len[(size_t)b]; });
   h->nsell = (int)perm.size();
   h->nstream = (int)streamrows.size();
   h->nlong = (int)longrows.size();
   perm.insert(perm.end(), streamrows.begin(), streamrows.end());
   perm.insert(perm.end(), longrows.begin(), longrows.end());
   const int nsx = h->nsell + h->nstream;

   std::vector<int> plen((size_t)nrows);
   for( int64_t i = 0; i < nrows; ++i )
      plen[(size_t)i] = len[(size_t)perm[(size_t)i]];

   // SELL slices
   const int nslices = (h->nsell + 31) / 32;
   std::vector<long long> sell_off((size_t)nslices + 1, 0);
   for( int sl = 0; sl < nslices; ++sl )
   {
      const int last = std::min(h->nsell, 32 * sl + 32) - 1;
      sell_off[(size_t)sl + 1] = sell_off[(size_t)sl] + 32LL * plen[(size_t)last];
   }
   // stream and long rows
   std::vector<long long> rowbeg((size_t)nrows + 1, 0);
   long long off = (sell_off[(size_t)nslices] + TILE - 1) / TILE * TILE;
   const long long streambase = off;
   for( int64_t i = h->nsell; i < nsx; ++i )
   {
      rowbeg[(size_t)i] = off;
      off += plen[(size_t)i];
   }
   h->nstreamelems = off - streambase;
   h->ntiles = (int)((h->nstreamelems + TILE - 1) / TILE);
   off = streambase + (long long)h->ntiles * TILE;      // the stream is padded with zero coefficients to a whole tile
   for( int64_t i = nsx; i < nrows; ++i )
   {
      off = (off + 3) & ~3LL;                           // long rows start 16-byte aligned
      rowbeg[(size_t)i] = off;
      off += plen[(size_t)i];
   }
   h->nstored = off;
   if( h->nstored >= (1LL << 40) || sell_off[(size_t)nslices] >= (1LL << 36) || h->nstreamelems / TILE >= (1LL << 31) - 64 || nrows >= (1LL << 30) )
   {
      gpulin_destroy(h);
      return fail(GPULIN_ERR_ARG, "matrix too large");
   }
   std::vector<double> pvals((size_t)h->nstored + 8, 0.0);
   std::vector<int> pcols((size_t)h->nstored + 8, 0);
   for( int64_t i = 0; i < nrows; ++i )
   {
      const int64_t r = perm[(size_t)i];
      const int64_t b = rowptr[r];
      const long long base = i < h->nsell ? sell_off[(size_t)(i >> 5)] + (i & 31) : rowbeg[(size_t)i];
      const long long stride = i < h->nsell ? 32 : 1;
      for( int k = 0; k < plen[(size_t)i]; ++k )
      {
         const int j = colidx[b + k];
         pvals[(size_t)(base + stride * k)] = vals[b + k];
         pcols[(size_t)(base + stride * k)] = j | (vartype[j] != 0 ? COL_INTEGRAL : 0) | (vals[b + k] < 0.0 ? COL_NEGCOEF : 0);
      }
   }
   std::vector<double2> sides((size_t)nrows + 1);
   for( int64_t i = 0; i < nrows; ++i )
      sides[(size_t)i] = make_double2(lhs[perm[(size_t)i]], rhs[perm[(size_t)i]]);

   // ---- rows the ranged-row rule looks at (rangedRowPropagation :5771-5776): kept on the host until gpulin_set_rangedrow
   h->rr_ptr.assign(1, 0);
   for( int64_t i = 0; i < nrows; ++i )
   {
      const int64_t r = perm[(size_t)i];
      if( plen[(size_t)i] >= 3 && lhs[r] > -num->infinity && rhs[r] < num->infinity )
      {
         h->rr_row.push_back((int)i);
         for( int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k )
         {
            h->rr_col.push_back(colidx[k]);
            h->rr_val.push_back(vals[k]);
         }
         h->rr_ptr.push_back((long long)h->rr_col.size());
      }
   }
   h->h_vartype.assign(vartype, vartype + ncols);

   // ---- tiles of the stream: row-end bit per nonzero, first unfinished row per tile ----------------------------------
   std::vector<unsigned char> endmask((size_t)h->ntiles * 32 + 32, 0);
   std::vector<int> tile_row0((size_t)h->ntiles + 2, nsx);
   {
      int t = 0;
      for( int i = h->nsell; i < nsx; ++i )
      {
         const long long last = rowbeg[(size_t)i] - streambase + plen[(size_t)i] - 1;
         endmask[(size_t)(last >> 3)] |= (unsigned char)(1u << (last & 7));
         // row i is the first unfinished row of every tile that starts at or before its last nonzero
         while( t < h->ntiles && (long long)t * TILE <= last )
            tile_row0[(size_t)t++] = i;
      }
   }

   // ---- column -> (permuted) rows --------------------------------------------------------------------------------
   std::vector<long long> colbeg((size_t)ncols + 2, 0);
   for( int64_t k = 0; k < nnz; ++k )
      ++colbeg[(size_t)colidx[k] + 1];
   for( int64_t j = 0; j < ncols; ++j )
      colbeg[(size_t)j + 1] += colbeg[(size_t)j];
   std::vector<int> colrows((size_t)nnz + 1);
   {
      std::vector<long long> fill(colbeg.begin(), colbeg.begin() + (size_t)ncols + 1);
      for( int64_t i = 0; i < nrows; ++i )
      {
         const int64_t r = perm[(size_t)i];
         for( int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k )
         {
            // flags: a tightened lower / upper bound of the column can matter for a finite side of this row
            const double a = vals[k];
            const bool rhsfin = rhs[r] < num->infinity;
            const bool lhsfin = lhs[r] > -num->infinity;
            int e = (int)i;
            if( (a > 0.0 && rhsfin) || (a < 0.0 && lhsfin) )
               e |= COLROW_LB;
            if( (a > 0.0 && lhsfin) || (a < 0.0 && rhsfin) )
               e |= COLROW_UB;
            colrows[(size_t)fill[(size_t)colidx[k]]++] = e;
         }
      }
   }

   // ---- upload ---------------------------------------------------------------------------------------------------
   DevProblem& p = h->p;
   long long* d_sell_off; int* d_sell_off32; unsigned char* d_coltype; unsigned char* d_colstate; int* d_flist; FastAcc* d_facc; unsigned* d_fracflag; int* d_rowlen; long long* d_rowbeg; double* d_vals; int* d_cols; double2* d_sides;
   int* d_tile_row0; unsigned char* d_endmask; unsigned char* d_tileflag;
   unsigned* d_freebits; double2* d_bndf;
   int* d_xlist; int* d_marklist; unsigned char* d_dirty; double2* d_bnd; long long* d_cand; unsigned* d_colbits; int* d_chglist; long long* d_colbeg; int* d_colrows;
   Ctrl* d_ctrl;
   std::vector<int> rowflags;      // rowlen words with their flags
   int rc = GPULIN_OK;
#define TRY(x) do { if( rc == GPULIN_OK ) rc = (x); } while( 0 )
#define TRYCU(x) do { if( rc == GPULIN_OK ) { cudaError_t e_ = (x); if( e_ != cudaSuccess ) rc = fail(GPULIN_ERR_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); } } while( 0 )
   TRY(devAlloc(h, &d_sell_off, (size_t)nslices + 1, true));
   TRY(devAlloc(h, &d_sell_off32, (size_t)nslices + 1, true));
   TRY(devAlloc(h, &d_coltype, (size_t)ncols + 1, true));
   TRY(devAlloc(h, &d_colstate, (size_t)ncols + 1));
   TRY(devAlloc(h, &d_flist, (size_t)h->nsell + 1));
   TRY(devAlloc(h, &d_facc, (size_t)h->nsell + 1));
   TRY(devAlloc(h, &d_fracflag, 1));
   TRY(devAlloc(h, &d_tile_row0, (size_t)h->ntiles + 2, true));
   TRY(devAlloc(h, &d_endmask, (size_t)h->ntiles * 32 + 32, true));
   TRY(devAlloc(h, &d_tileflag, (size_t)h->ntiles + 64));
   TRY(devAlloc(h, &d_rowlen, (size_t)nrows + 1, true));
   TRY(devAlloc(h, &d_rowbeg, (size_t)nrows + 1, true));
   TRY(devAlloc(h, &d_vals, (size_t)h->nstored + 8, true));
   TRY(devAlloc(h, &d_cols, (size_t)h->nstored + 8, true));
   TRY(devAlloc(h, &d_sides, (size_t)nrows + 1, true));
   TRY(devAlloc(h, &d_dirty, (size_t)nrows + 64));
   TRY(devAlloc(h, &d_xlist, (size_t)nrows + 1));
   TRY(devAlloc(h, &d_marklist, (size_t)6 * MARKCAP));
   TRY(devAlloc(h, &d_bnd, (size_t)ncols + 1));
   TRY(devAlloc(h, &d_cand, 2 * (size_t)ncols + 2));
   TRY(devAlloc(h, &d_colbits, (size_t)ncols / 32 + 2));
   const int nallwords = (int)((((size_t)ncols + 31) / 32 + 3) & ~(size_t)3);
   h->freebytes = sizeof(unsigned) * ((size_t)nallwords + 4);
   TRY(devAlloc(h, &d_freebits, (size_t)nallwords + 4));
   TRY(devAlloc(h, &d_bndf, (size_t)ncols + 1));
   TRY(devAlloc(h, &d_chglist, (size_t)ncols + 1));
   TRY(devAlloc(h, &d_colbeg, (size_t)ncols + 2, true));
   TRY(devAlloc(h, &d_colrows, (size_t)nnz + 1, true));
   TRY(devAlloc(h, &d_ctrl, 1));
   TRY(devAlloc(h, &h->d_peers, 1));
   TRY(devAlloc(h, &h->d_tmplb, (size_t)ncols + 1));
   TRY(devAlloc(h, &h->d_tmpub, (size_t)ncols + 1));
   TRYCU(cudaMemcpy(d_sell_off, sell_off.data(), sizeof(long long) * ((size_t)nslices + 1), cudaMemcpyHostToDevice));
   {
      std::vector<int> off32((size_t)nslices + 1);
      for( int sl = 0; sl <= nslices; ++sl )
         off32[(size_t)sl] = (int)(sell_off[(size_t)sl] >> 5);
      TRYCU(cudaMemcpy(d_sell_off32, off32.data(), sizeof(int) * ((size_t)nslices + 1), cudaMemcpyHostToDevice));
   }
   TRYCU(cudaMemcpy(d_coltype, vartype, (size_t)ncols, cudaMemcpyHostToDevice));
   TRYCU(cudaMemset(d_colstate, CS_OTHER, (size_t)ncols + 1));
   TRYCU(cudaMemset(d_fracflag, 0, sizeof(unsigned)));
   TRYCU(cudaMemcpy(d_tile_row0, tile_row0.data(), sizeof(int) * ((size_t)h->ntiles + 1), cudaMemcpyHostToDevice));
   TRYCU(cudaMemcpy(d_endmask, endmask.data(), (size_t)h->ntiles * 32, cudaMemcpyHostToDevice));
   TRYCU(cudaMemset(d_tileflag, 0, (size_t)h->ntiles + 64));
   {
      // rows with a coefficient below hugeval / infinity always take the exact rules (see ROWLEN_EXACT)
      const double tiny = num->hugeval / num->infinity;
      std::vector<int>& flagged = rowflags;
      flagged = plen;
      for( int64_t i = 0; i < nrows; ++i )
      {
         const int64_t r = perm[(size_t)i];
         for( int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k )
         {
            if( std::fabs(vals[k]) < tiny )
            {
               flagged[(size_t)i] |= ROWLEN_EXACT;
               break;
            }
         }
         // integer coefficients over columns of integral type: the filter's sums can be exact (see FastAcc)
         bool allint = true;
         for( int64_t k = rowptr[r]; k < rowptr[r + 1] && allint; ++k )
            allint = vartype[colidx[k]] != 0 && std::fabs(vals[k]) <= 1048576.0 && vals[k] == std::rint(vals[k]);
         if( allint )
            flagged[(size_t)i] |= ROWLEN_INT;
      }
      TRYCU(cudaMemcpy(d_rowlen, flagged.data(), sizeof(int) * (size_t)nrows, cudaMemcpyHostToDevice));
   }
   TRYCU(cudaMemcpy(d_rowbeg, rowbeg.data(), sizeof(long long) * ((size_t)nrows + 1), cudaMemcpyHostToDevice));
   TRYCU(cudaMemcpy(d_vals, pvals.data(), sizeof(double) * ((size_t)h->nstored + 8), cudaMemcpyHostToDevice));
   TRYCU(cudaMemcpy(d_cols, pcols.data(), sizeof(int) * ((size_t)h->nstored + 8), cudaMemcpyHostToDevice));
   TRYCU(cudaMemcpy(d_sides, sides.data(), sizeof(double2) * (size_t)nrows, cudaMemcpyHostToDevice));
   TRYCU(cudaMemcpy(d_colbeg, colbeg.data(), sizeof(long long) * ((size_t)ncols + 1), cudaMemcpyHostToDevice));
   TRYCU(cudaMemcpy(d_colrows, colrows.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
   TRYCU(cudaMemset(d_dirty, 0, (size_t)nrows + 64));
   TRYCU(cudaMemset(d_colbits, 0, sizeof(unsigned) * ((size_t)ncols / 32 + 2)));
   TRYCU(cudaMemset(d_ctrl, 0, sizeof(Ctrl)));
   TRYCU(cudaMemset(d_cand, 0, sizeof(long long) * (2 * (size_t)ncols + 2)));
   TRYCU(cudaMallocHost((void**)&h->h_ctrl, sizeof(Ctrl)));
   TRYCU(cudaMallocHost((void**)&h->h_params, 4 * sizeof(int)));
   if( rc == GPULIN_OK )
      memset(h->h_ctrl, 0, sizeof(Ctrl));
   TRYCU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
   TRYCU(cudaStreamCreateWithFlags(&h->aux[0], cudaStreamNonBlocking));
   TRYCU(cudaStreamCreateWithFlags(&h->aux[1], cudaStreamNonBlocking));
   TRYCU(cudaEventCreateWithFlags(&h->evfork, cudaEventDisableTiming));
   TRYCU(cudaEventCreateWithFlags(&h->evjoin[0], cudaEventDisableTiming));
   TRYCU(cudaEventCreateWithFlags(&h->evjoin[1], cudaEventDisableTiming));
   TRYCU(cudaEventCreateWithFlags(&h->evfork2, cudaEventDisableTiming));
   TRYCU(cudaEventCreateWithFlags(&h->evjoin2, cudaEventDisableTiming));
   TRYCU(cudaEventCreate(&h->ev0));
   TRYCU(cudaEventCreate(&h->ev1));
   if( rc != GPULIN_OK )
   {
      gpulin_destroy(h);
      return rc;
   }

   p.nrows = (int)nrows;
   p.ncols = (int)ncols;
   p.nsell = h->nsell;
   p.nsellunit = h->nsellunit;
   p.nsx = nsx;
   p.ntiles = h->ntiles;
   p.streambase = streambase;
   p.sell_off = d_sell_off;
   p.sell_off32 = d_sell_off32;
   p.coltype = d_coltype;
   p.colstate = d_colstate;
   p.flist = d_flist;
   p.facc = d_facc;
   p.fracflag = d_fracflag;
   p.tile_row0 = d_tile_row0;
   p.endmask = d_endmask;
   p.tileflag = d_tileflag;
   p.rowlen = d_rowlen;
   p.rowbeg = d_rowbeg;
   p.vals = d_vals;
   p.cols = d_cols;
   p.sides = d_sides;
   p.dirty = d_dirty;
   p.xlist = d_xlist;
   p.marklist = d_marklist;
   p.bnd = d_bnd;
   p.freebits = d_freebits;
   p.bndf = d_bndf;
   p.nfreewords = std::min(nallwords, SELLBITS_MAXWORDS);
   p.nfreecols = (int)std::min<int64_t>(ncols, (int64_t)p.nfreewords * 32);
   p.cand = d_cand;
   p.colbits = d_colbits;
   p.chglist = d_chglist;
   p.colbeg = d_colbeg;
   p.colrows = d_colrows;
   p.ctrl = d_ctrl;
   p.log = nullptr;
   p.peers = nullptr;
   memset(&p.rr, 0, sizeof(p.rr));
   setShare(h, 0, 1);
   {
      // expected marks per row >= 1: marking every row is cheaper than walking the columns (see apply_kernel) -- where a
      // spurious mark is cheap: the filter of the thread-per-row rows costs ~3 ps per nonzero and finishes almost every
      // row it did not have to look at, a longer row that the filter cannot finish costs a pass of the exact rules (and
      // the marking rule spares most rows of a knapsack-type matrix: an upper bound that moves down does not concern
      // a <= row with positive coefficients).  So only for matrices whose nonzeros are in short rows.  (Tried on C4, where
      // the first round changes 677k columns of 25 rows each: marking everything takes 114 us off its apply step and puts
      // 330 us on the second round, which sweeps and re-examines four times the rows.)
      long long sellnnz = 0;
      for( int i = 0; i < h->nsell; ++i )
         sellnnz += plen[(size_t)i];
      const double thr = nnz > 0 ? std::ceil((double)nrows * (double)ncols / (double)nnz) : 4e9;
      p.markall_min = (sellnnz >= (nnz / 10) * 9) ? (unsigned)std::min(thr, 4e9) : 0xffffffffu;
   }
   p.num.inf = num->infinity;
   p.num.eps = num->epsilon;
   p.num.sumeps = num->sumepsilon;
   p.num.feastol = num->feastol;
   p.num.bstreps = num->boundstreps;
   p.num.huge = num->hugeval;
   p.num.maxeasy = num->maxeasyactivitydelta;

   // ---- launch geometry ------------------------------------------------------------------------------------------
   // persistent grids: as many blocks as stay resident, never more than there is work
   {
      int occa = 0;
      if( cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occa, apply_kernel<APPLY_LIST, true>, APPLY_THREADS, 0) != cudaSuccess || occa < 1 )
         occa = 2;
      h->napplyblocks = (int)std::max<int64_t>(1, std::min<int64_t>((ncols + APPLY_THREADS - 1) / APPLY_THREADS, (int64_t)h->nsm * occa));
   }
   {
      int occ = 0;
      if( cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_stream_kernel, SWEEP_THREADS, 0) != cudaSuccess || occ < 1 )
         occ = 1;
      const int wpb = SWEEP_THREADS / 32;
      // at least two tiles per warp, never more blocks than stay resident
      const int64_t need = ((int64_t)h->ntiles + 2 * wpb - 1) / (2 * wpb);
      h->nstreamblocks = (int)std::min<int64_t>(need, (int64_t)h->nsm * occ);
      int occs = 0;
      if( cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occs, g_sellKernel, SELL_THREADS, 0) != cudaSuccess || occs < 1 )
         occs = 1;
      const int64_t needs = ((int64_t)nslices + (SELL_THREADS / 32) - 1) / (SELL_THREADS / 32);
      h->nsellblocks = (int)std::min<int64_t>(needs, (int64_t)h->nsm * occs);
      // the shared-memory bit-table variant: one block of 1024 threads per SM; worth its prologue (the table is staged
      // by every block) once there is more than a handful of slices per block
      h->sellbitsvariant = h->p.nfreecols >= h->p.ncols ? 1 : 0;
      if( h->nsellblocks > 0 && nslices >= h->nsm * 32
         && cudaFuncSetAttribute(g_sellBitsKernels[h->sellbitsvariant], cudaFuncAttributeMaxDynamicSharedMemorySize,
               (int)sellBitsSmem(SELLBITS_MAXWORDS)) == cudaSuccess )
         h->nsellbitsblocks = h->nsm;
      if( cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_long_kernel, LONG_THREADS, 0) != cudaSuccess || occ < 1 )
         occ = 1;
      h->nlongblocks = (int)std::min<int64_t>(h->nlong, (int64_t)h->nsm * occ);
      // two resident blocks per SM with 128 registers
      h->exactkernel = exact_rows_kernel<2>;
      if( cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->exactkernel, EXACT_THREADS, 0) != cudaSuccess || occ < 1 )
         occ = 1;
      h->nexactblocks = (int)std::max<int64_t>(1, std::min<int64_t>((nrows + EXACT_THREADS - 1) / EXACT_THREADS, (int64_t)h->nsm * occ));
      // fast_rows_kernel (a thread per row) takes the rows that came with their activities when there are enough of them:
      // it needs ~36 us whatever their number (up to 32 rows per resident warp), eight lanes per row in exact_rows_kernel
      // need ~6.4 us per trip of 4 rows per warp -- the break-even is at ~22 rows per warp (measured on C3;
      // GPULIN_FASTMIN overrides the threshold: the tests run the kernel on small instances with it).  Launched only if
      // the thread-per-row class holds a row that can qualify (ROWLEN_INT).
      {
         int occf = 0;
         if( cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occf, fast_rows_kernel, FAST_THREADS, 0) != cudaSuccess || occf < 1 )
            occf = 1;
         bool anyint = false;
         for( int64_t i = 0; i < h->nsell && !anyint; ++i )
            anyint = (rowflags[(size_t)i] & ROWLEN_INT) != 0;
         h->nfastblocks = anyint ? (int)std::max<int64_t>(1, std::min<int64_t>((h->nsell + FAST_THREADS - 1) / FAST_THREADS, (int64_t)h->nsm * occf)) : 0;
      }
      p.fastmin = 22u * (unsigned)h->nexactblocks * (EXACT_THREADS / 32);
      if( getenv("GPULIN_FASTMIN") != nullptr )
         p.fastmin = (unsigned)std::max(0, atoi(getenv("GPULIN_FASTMIN")));
      h->npushblocks = (int)std::max<int64_t>(1, std::min<int64_t>((ncols / 32 + 255) / 256, (int64_t)h->nsm * 4));
      // the sparse-rounds kernel: one block per SM (all must be co-resident: grid syncs)
      int coop = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
      if( coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sparse_rounds_kernel<true>, SPARSE_THREADS, 0) == cudaSuccess && occ >= 1 )
         h->nsparseblocks = h->nsm;
   }

   if( !h->hostloop )
   {
      rc = buildGraph(h);
      if( rc != GPULIN_OK )
      {
         gpulin_destroy(h);
         return rc;
      }
   }
   *out = h;
   return GPULIN_OK;
}

extern "C" void gpulin_destroy(gpulin_t* h)
{
   if( h == nullptr )
      return;
   cudaSetDevice(h->device);
   for( gpulin* w : h->workers )
      gpulin_destroy(w);
   h->workers.clear();
   cudaFree(h->d_probelog);
   cudaFree(h->d_probecursor);
   cudaFree(h->d_proberes);
   cudaFree(h->d_probevar);
   cudaFree(h->d_probelb);
   cudaFree(h->d_probeub);
   h->d_proberes = nullptr;
   h->d_probevar = nullptr;
   h->d_probelb = h->d_probeub = nullptr;
   if( h->stream != nullptr )
      cudaStreamSynchronize(h->stream);
   destroyGraph(h);
   for( int r = 0; r < MAX_PEERS; ++r )
   {
      if( h->peerptr[r] != nullptr )
         cudaIpcCloseMemHandle(h->peerptr[r]);
   }
   cudaFree(h->d_inbox);
   for( int i = 0; i < h->nalloc; ++i )
      cudaFree(h->d_all[i]);
   if( h->shared != nullptr && --h->shared->refs == 0 )
   {
      for( int i = 0; i < h->shared->n; ++i )
         cudaFree(h->shared->ptrs[i]);
      delete h->shared;
   }
   if( h->d_log != nullptr )
      cudaFree(h->d_log);
   if( h->h_ctrl != nullptr )
      cudaFreeHost(h->h_ctrl);
   if( h->h_params != nullptr )
      cudaFreeHost(h->h_params);
   cudaFree(h->d_redflags);
   if( h->h_redflags != nullptr )
      cudaFreeHost(h->h_redflags);
   cudaFree(h->d_updidx);
   cudaFree(h->d_updlb);
   cudaFree(h->d_updub);
   cudaFree(h->d_rr[5]);
   cudaFree(h->d_ref);
   cudaFree(h->d_codes);
   cudaFree(h->d_packlog);
   cudaFree(h->d_clog);
   if( h->h_clogcount != nullptr )
      cudaFreeHost(h->h_clogcount);
   for( int i = 0; i < 2; ++i )
   {
      if( h->aux[i] != nullptr )
         cudaStreamDestroy(h->aux[i]);
      if( h->evjoin[i] != nullptr )
         cudaEventDestroy(h->evjoin[i]);
   }
   for( int i = 0; i < 4; ++i )
   {
      if( h->evprof[i] != nullptr )
         cudaEventDestroy(h->evprof[i]);
   }
   if( h->evfork != nullptr )
      cudaEventDestroy(h->evfork);
   if( h->evfork2 != nullptr )
      cudaEventDestroy(h->evfork2);
   if( h->evjoin2 != nullptr )
      cudaEventDestroy(h->evjoin2);
   if( h->ev0 != nullptr )
      cudaEventDestroy(h->ev0);
   if( h->ev1 != nullptr )
      cudaEventDestroy(h->ev1);
   if( h->stream != nullptr && h->ownstream )
      cudaStreamDestroy(h->stream);
   delete h;
}

static int gridFor(const gpulin* h, int64_t n)
{
   return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)h->nsm * 8));
}

extern "C" int gpulin_set_stream(gpulin_t* h, void* stream)
{
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   if( stream == nullptr && !h->hostloop )
      return fail(GPULIN_ERR_ARG, "the legacy default stream cannot be captured into the round graph: pass a created stream");
   CU(cudaSetDevice(h->device));
   CU(cudaStreamSynchronize(h->stream));
   if( h->ownstream && h->stream != nullptr )
      cudaStreamDestroy(h->stream);
   h->stream = (cudaStream_t)stream;
   h->ownstream = false;
   if( !h->hostloop )
      OK(buildGraph(h));
   return GPULIN_OK;
}

extern "C" int gpulin_set_bounds_device(gpulin_t* h, const double* d_lb, const double* d_ub)
{
   if( h == nullptr || d_lb == nullptr || d_ub == nullptr )
      return fail(GPULIN_ERR_ARG, "NULL argument");
   CU(cudaSetDevice(h->device));
   CU(cudaMemsetAsync(h->p.fracflag, 0, sizeof(unsigned), h->stream));
   set_bounds_kernel<<<gridFor(h, std::max(h->ncols, h->nrows)), 256, 0, h->stream>>>(h->p, d_lb, d_ub);
   h->smallcols = -1;
   ++h->version;
   CU(cudaGetLastError());
   h->havebounds = true;
   return GPULIN_OK;
}

extern "C" int gpulin_set_bounds(gpulin_t* h, const double* lb, const double* ub)
{
   if( h == nullptr || lb == nullptr || ub == nullptr )
      return fail(GPULIN_ERR_ARG, "NULL argument");
   CU(cudaSetDevice(h->device));
   CU(cudaMemcpyAsync(h->d_tmplb, lb, sizeof(double) * (size_t)h->ncols, cudaMemcpyHostToDevice, h->stream));
   CU(cudaMemcpyAsync(h->d_tmpub, ub, sizeof(double) * (size_t)h->ncols, cudaMemcpyHostToDevice, h->stream));
   return gpulin_set_bounds_device(h, h->d_tmplb, h->d_tmpub);
}

static int growUpdateStaging(gpulin* h, int64_t n)
{
   if( n <= h->updcap )
      return GPULIN_OK;
   CU(cudaStreamSynchronize(h->stream));
   cudaFree(h->d_updidx);
   cudaFree(h->d_updlb);
   cudaFree(h->d_updub);
   h->d_updidx = nullptr;
   h->d_updlb = nullptr;
   h->d_updub = nullptr;
   h->updcap = 0;
   const int64_t cap = std::max<int64_t>(n, 1024);
   if( cudaMalloc((void**)&h->d_updidx, sizeof(int) * (size_t)cap) != cudaSuccess
      || cudaMalloc((void**)&h->d_updlb, sizeof(double) * (size_t)cap) != cudaSuccess
      || cudaMalloc((void**)&h->d_updub, sizeof(double) * (size_t)cap) != cudaSuccess )
      return fail(GPULIN_ERR_NOMEM, "cudaMalloc of the bound-update staging failed");
   h->updcap = cap;
   return GPULIN_OK;
}

extern "C" int gpulin_set_reference_bounds(gpulin_t* h, const double* lb, const double* ub)
{
   if( h == nullptr || lb == nullptr || ub == nullptr )
      return fail(GPULIN_ERR_ARG, "NULL argument");
   CU(cudaSetDevice(h->device));
   if( h->d_ref == nullptr )
   {
      if( cudaMalloc((void**)&h->d_ref, sizeof(double2) * ((size_t)h->ncols + 1)) != cudaSuccess
         || cudaMalloc((void**)&h->d_codes, sizeof(unsigned) * ((size_t)h->ncols / 16 + 2)) != cudaSuccess )
         return fail(GPULIN_ERR_NOMEM, "cudaMalloc of the reference bounds failed");
   }
   CU(cudaMemcpyAsync(h->d_tmplb, lb, sizeof(double) * (size_t)h->ncols, cudaMemcpyHostToDevice, h->stream));
   CU(cudaMemcpyAsync(h->d_tmpub, ub, sizeof(double) * (size_t)h->ncols, cudaMemcpyHostToDevice, h->stream));
   set_reference_kernel<<<gridFor(h, h->ncols), 256, 0, h->stream>>>((int)h->ncols, h->d_tmplb, h->d_tmpub, h->d_ref);
   CU(cudaGetLastError());
   CU(cudaStreamSynchronize(h->stream));     // (the host arrays may be pageable and temporary)
   return GPULIN_OK;
}

extern "C" int gpulin_set_bounds_packed(gpulin_t* h, const uint32_t* codes, int64_t nexplicit, const int32_t* idx,
   const double* lb, const double* ub)
{
   if( h == nullptr || codes == nullptr || nexplicit < 0 || (nexplicit > 0 && (idx == nullptr || lb == nullptr || ub == nullptr)) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( h->d_ref == nullptr )
      return fail(GPULIN_ERR_STATE, "gpulin_set_bounds_packed before gpulin_set_reference_bounds");
   for( int64_t i = 0; i < nexplicit; ++i )
   {
      if( idx[i] < 0 || idx[i] >= h->ncols )
         return fail(GPULIN_ERR_ARG, "column index %d out of range", idx[i]);
   }
   CU(cudaSetDevice(h->device));
   const size_t nwords = ((size_t)h->ncols + 15) / 16;
   CU(cudaMemcpyAsync(h->d_codes, codes, sizeof(unsigned) * nwords, cudaMemcpyHostToDevice, h->stream));
   CU(cudaMemsetAsync(h->p.fracflag, 0, sizeof(unsigned), h->stream));
   set_bounds_packed_kernel<<<gridFor(h, std::max(h->ncols, h->nrows)), 256, 0, h->stream>>>(h->p, h->d_ref, h->d_codes);
   if( nexplicit > 0 )
   {
      OK(growUpdateStaging(h, nexplicit));
      CU(cudaMemcpyAsync(h->d_updidx, idx, sizeof(int) * (size_t)nexplicit, cudaMemcpyHostToDevice, h->stream));
      CU(cudaMemcpyAsync(h->d_updlb, lb, sizeof(double) * (size_t)nexplicit, cudaMemcpyHostToDevice, h->stream));
      CU(cudaMemcpyAsync(h->d_updub, ub, sizeof(double) * (size_t)nexplicit, cudaMemcpyHostToDevice, h->stream));
      update_explicit_kernel<<<gridFor(h, nexplicit), 256, 0, h->stream>>>(h->p, nexplicit, h->d_updidx, h->d_updlb, h->d_updub);
   }
   CU(cudaGetLastError());
   h->smallcols = -1;
   ++h->version;
   h->havebounds = true;
   return GPULIN_OK;
}

extern "C" int gpulin_get_changes_packed(gpulin_t* h, void* out, int64_t maxn, int64_t* n)
{
   if( h == nullptr || n == nullptr || maxn < 0 || (maxn > 0 && out == nullptr) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   CU(cudaSetDevice(h->device));
   const int64_t produced = (int64_t)h->h_ctrl->logcount;
   *n = produced;
   const int64_t m = std::min(std::min(produced, h->logcap), maxn);
   if( m <= 0 )
      return GPULIN_OK;
   if( m > h->packlogcap )
   {
      CU(cudaStreamSynchronize(h->stream));
      cudaFree(h->d_packlog);
      h->d_packlog = nullptr;
      h->packlogcap = 0;
      const int64_t cap = std::max<int64_t>(h->logcap, m);
      if( cudaMalloc((void**)&h->d_packlog, 12 * (size_t)cap) != cudaSuccess )
         return fail(GPULIN_ERR_NOMEM, "cudaMalloc of the packed change log failed");
      h->packlogcap = cap;
   }
   pack_log_kernel<<<gridFor(h, m), 256, 0, h->stream>>>(h->d_log, m, h->d_packlog);
   CU(cudaGetLastError());
   CU(cudaMemcpyAsync(out, h->d_packlog, 12 * (size_t)m, cudaMemcpyDeviceToHost, h->stream));
   CU(cudaStreamSynchronize(h->stream));
   return GPULIN_OK;
}

extern "C" int gpulin_get_changes_compact(gpulin_t* h, uint32_t* out, int64_t maxn, int64_t* n, uint32_t* xout, int64_t maxx,
   int64_t* nx)
{
   if( h == nullptr || n == nullptr || nx == nullptr || maxn < 0 || maxx < 0 || (maxn > 0 && out == nullptr) || (maxx > 0 && xout == nullptr) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( h->ncols >= (1LL << 29) )
      return fail(GPULIN_ERR_ARG, "the compact change log holds column indices below 2^29");
   CU(cudaSetDevice(h->device));
   const int64_t produced = (int64_t)h->h_ctrl->logcount;
   *n = produced;
   *nx = 0;
   const int64_t m = std::min(std::min(produced, h->logcap), maxn);
   if( m <= 0 )
      return GPULIN_OK;
   if( m >= (1LL << 32) )
      return fail(GPULIN_ERR_ARG, "more than 2^32 log entries");
   if( m > h->clogcap )
   {
      CU(cudaStreamSynchronize(h->stream));
      cudaFree(h->d_clog);
      h->d_clog = nullptr;
      h->clogcap = 0;
      const int64_t cap = std::max<int64_t>(h->logcap, m);
      // [ explicit count (16 bytes) | one word per entry | three words per explicit entry ]
      if( cudaMalloc((void**)&h->d_clog, 16 + 16 * (size_t)cap) != cudaSuccess )
         return fail(GPULIN_ERR_NOMEM, "cudaMalloc of the compact change log failed");
      h->clogcap = cap;
   }
   if( h->h_clogcount == nullptr )
      CU(cudaMallocHost((void**)&h->h_clogcount, sizeof(unsigned)));
   unsigned* d_count = h->d_clog;
   unsigned* d_words = h->d_clog + 4;
   unsigned* d_x = d_words + h->clogcap;
   CU(cudaMemsetAsync(d_count, 0, sizeof(unsigned), h->stream));
   compact_log_kernel<<<gridFor(h, m), 256, 0, h->stream>>>(h->d_log, m, d_words, d_x, d_count);
   CU(cudaGetLastError());
   CU(cudaMemcpyAsync(out, d_words, 4 * (size_t)m, cudaMemcpyDeviceToHost, h->stream));
   CU(cudaMemcpyAsync(h->h_clogcount, d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
   CU(cudaStreamSynchronize(h->stream));
   *nx = (int64_t)*h->h_clogcount;
   const int64_t mx = std::min(*nx, maxx);
   if( mx > 0 )
   {
      CU(cudaMemcpyAsync(xout, d_x, 12 * (size_t)mx, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
   }
   return GPULIN_OK;
}

// ranged-row propagation on / off.  On: every ranged row is copied in the order the reference walks it -- the order of
// consdataCompVarProp (cons_linear.c:3191-3257): binaries by decreasing |a|, the other integers by decreasing
// |a (ub_global - lb_global)|, continuous variables; ties by SCIPvarGetProbindex (`tie`, NULL: the column index)
extern "C" int gpulin_set_rangedrow(gpulin_t* h, int enable, const double* glb, const double* gub, const int32_t* tie)
{
   if( h == nullptr || (enable && (glb == nullptr || gub == nullptr)) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( h->npeers > 1 && (h->p.rr.n > 0) != (enable != 0 && !h->rr_row.empty()) )
      return fail(GPULIN_ERR_STATE, "gpulin_set_rangedrow must be called before the handle is connected to peers");
   CU(cudaSetDevice(h->device));
   CU(cudaStreamSynchronize(h->stream));
   // (the row copies belong to the shared matrix -- clones use them; arrays of an earlier call stay until it goes --,
   // the scratch words to this handle)
   cudaFree(h->d_rr[5]);
   for( int i = 0; i < 6; ++i )
      h->d_rr[i] = nullptr;
   memset(&h->p.rr, 0, sizeof(h->p.rr));
   const int n = (int)h->rr_row.size();
   if( enable && n > 0 && h->shared->n + 5 > 64 )
      return fail(GPULIN_ERR_STATE, "gpulin_set_rangedrow called too often on this matrix");
   if( enable && n > 0 )
   {
      const size_t nnzr = h->rr_col.size();
      std::vector<int> scols(nnzr);
      std::vector<double> svals(nnzr);
      std::vector<int> order;
      int maxlen = 0;
      auto isbin = [&](int j) { return h->h_vartype[(size_t)j] != 0 && glb[j] >= 0.0 && gub[j] <= 1.0; };
      auto tiekey = [&](int j) { return tie != nullptr ? (long long)tie[j] : (long long)j; };
      for( int i = 0; i < n; ++i )
      {
         const long long b = h->rr_ptr[(size_t)i];
         const int len = (int)(h->rr_ptr[(size_t)i + 1] - b);
         maxlen = std::max(maxlen, len);
         order.resize((size_t)len);
         std::iota(order.begin(), order.end(), 0);
         const int* c = h->rr_col.data() + b;
         const double* a = h->rr_val.data() + b;
         std::sort(order.begin(), order.end(), [&](int x, int y) {
            const int j1 = c[x];
            const int j2 = c[y];
            const bool b1 = isbin(j1);
            const bool b2 = isbin(j2);
            if( b1 != b2 )
               return b1;
            if( b1 )
            {
               const double a1 = std::fabs(a[x]);
               const double a2 = std::fabs(a[y]);
               if( a1 - a2 > 1e-9 ) return true;
               if( a2 - a1 > 1e-9 ) return false;
               return tiekey(j1) < tiekey(j2);
            }
            const bool i1 = h->h_vartype[(size_t)j1] != 0;
            const bool i2 = h->h_vartype[(size_t)j2] != 0;
            if( i1 != i2 )
               return i1;
            if( !i1 )
               return tiekey(j1) < tiekey(j2);
            const double c1 = std::fabs(a[x] * (gub[j1] - glb[j1]));
            const double c2 = std::fabs(a[y] * (gub[j2] - glb[j2]));
            if( c1 - c2 > 1e-9 ) return true;
            if( c2 - c1 > 1e-9 ) return false;
            return tiekey(j1) < tiekey(j2);
         });
         for( int v = 0; v < len; ++v )
         {
            const int j = c[order[(size_t)v]];
            scols[(size_t)(b + v)] = j | (h->h_vartype[(size_t)j] != 0 ? (int)0x80000000u : 0);
            svals[(size_t)(b + v)] = a[order[(size_t)v]];
         }
      }
      std::vector<int> idx((size_t)h->nrows + 1, -1);
      for( int i = 0; i < n; ++i )
         idx[(size_t)h->rr_row[(size_t)i]] = i;
      const int scratchwords = (maxlen + 31) / 32 + 1;
      const size_t nwarps = (size_t)h->nsm * 64;      // more warps than any grid that runs the rule has
      size_t bytes[6] = {sizeof(long long) * ((size_t)n + 1), sizeof(int) * (size_t)n, sizeof(int) * nnzr, sizeof(double) * nnzr,
         sizeof(int) * ((size_t)h->nrows + 1), sizeof(unsigned) * nwarps * (size_t)scratchwords};
      const void* src[6] = {h->rr_ptr.data(), h->rr_row.data(), scols.data(), svals.data(), idx.data(), nullptr};
      for( int i = 0; i < 6; ++i )
      {
         if( cudaMalloc(&h->d_rr[i], std::max<size_t>(bytes[i], 16)) != cudaSuccess )
            return fail(GPULIN_ERR_NOMEM, "cudaMalloc of the ranged rows failed");
         if( src[i] != nullptr )
            CU(cudaMemcpy(h->d_rr[i], src[i], bytes[i], cudaMemcpyHostToDevice));
         if( i < 5 )
            h->shared->ptrs[h->shared->n++] = h->d_rr[i];
      }
      RangedRows& R = h->p.rr;
      R.n = n;
      R.beg = (const long long*)h->d_rr[0];
      R.row = (const int*)h->d_rr[1];
      R.cols = (const int*)h->d_rr[2];
      R.vals = (const double*)h->d_rr[3];
      R.idx = (const int*)h->d_rr[4];
      R.scratch = (unsigned*)h->d_rr[5];
      R.scratchwords = scratchwords;
      h->nrangedblocks = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)n * 32 + RANGED_THREADS - 1) / RANGED_THREADS, (int64_t)h->nsm * 8));
   }
   // kernel parameters are baked into the graph
   if( !h->hostloop )
      OK(buildGraph(h));
   return GPULIN_OK;
}

extern "C" int gpulin_update_bounds(gpulin_t* h, int64_t n, const int32_t* idx, const double* lb, const double* ub)
{
   if( h == nullptr || n < 0 || (n > 0 && (idx == nullptr || lb == nullptr || ub == nullptr)) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( !h->havebounds )
      return fail(GPULIN_ERR_STATE, "gpulin_update_bounds before gpulin_set_bounds");
   if( n == 0 )
      return GPULIN_OK;
   for( int64_t i = 0; i < n; ++i )
   {
      if( idx[i] < 0 || idx[i] >= h->ncols )
         return fail(GPULIN_ERR_ARG, "column index %d out of range", idx[i]);
   }
   CU(cudaSetDevice(h->device));
   ++h->version;
   if( n <= SmallUpdate::CAP )
   {
      // a handful of bounds (the usual call at a branch-and-bound node): passed by value, no staging copies
      SmallUpdate su;
      su.n = (int)n;
      for( int64_t i = 0; i < n; ++i )
      {
         su.idx[i] = idx[i];
         su.lb[i] = lb[i];
         su.ub[i] = ub[i];
      }
      update_small_kernel<<<1, 32 * SmallUpdate::CAP, 0, h->stream>>>(h->p, su);
      CU(cudaGetLastError());
      if( h->smallcols >= 0 )
         h->smallcols += n;
      return GPULIN_OK;
   }
   OK(growUpdateStaging(h, n));
   CU(cudaMemcpyAsync(h->d_updidx, idx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
   CU(cudaMemcpyAsync(h->d_updlb, lb, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
   CU(cudaMemcpyAsync(h->d_updub, ub, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
   update_bounds_kernel<<<gridFor(h, n), 256, 0, h->stream>>>(h->p, n, h->d_updidx, h->d_updlb, h->d_updub);
   CU(cudaGetLastError());
   if( h->smallcols >= 0 )
      h->smallcols += n;
   return GPULIN_OK;
}

static int fetchCtrl(gpulin* h)
{
   CU(cudaMemcpyAsync(h->h_ctrl, h->p.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
   CU(cudaStreamSynchronize(h->stream));
   return GPULIN_OK;
}

extern "C" int gpulin_propagate_async(gpulin_t* h, int maxrounds)
{
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   if( !h->havebounds )
      return fail(GPULIN_ERR_STATE, "gpulin_propagate before gpulin_set_bounds");
   CU(cudaSetDevice(h->device));
   h->lastmaxrounds = maxrounds;
   ++h->version;
   h->smallcall = h->smallcols >= 0 && h->smallcols <= SMALLCALL_MAXCOLS && !h->hostloop && h->smallcalls;
   h->smallcols = -1;
   h->h_params[0] = maxrounds;
   h->h_params[1] = (int)h->logcap;
   if( !h->smallcall )       // (probe_kernel takes both as arguments)
      CU(cudaMemcpyAsync(&h->p.ctrl->maxrounds, h->h_params, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
   CU(cudaEventRecord(h->ev0, h->stream));
   if( h->smallcall )
   {
      // few bounds moved since the last fixpoint: one block runs the whole call (and hands over if the cascade grows)
      probe_kernel<<<1, PROBE_THREADS, 0, h->stream>>>(h->p, h->p, -1, -1, 0.0, 0.0, maxrounds, (int)h->logcap, 1, nullptr);
      CU(cudaGetLastError());
   }
   else if( !h->hostloop )
   {
      CU(cudaGraphLaunch(h->gexec, h->stream));
   }
   else
   {
      begin_kernel<<<1, 1, 0, h->stream>>>(h->p.ctrl);
      for( bool first = true; ; first = false )
      {
         OK((launchRoundKernels<APPLY_LIST, false>(h, true, true, true, first)));
         int cont = 0;
         CU(cudaMemcpyAsync(&h->h_ctrl->cont, &h->p.ctrl->cont, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
         CU(cudaStreamSynchronize(h->stream));
         cont = h->h_ctrl->cont;
         if( !cont )
            break;
      }
   }
   CU(cudaEventRecord(h->ev1, h->stream));
   // the verdict and the statistics come back with the same stream order, without blocking the caller
   // (the per-round history is fetched when gpulin_get_round_stats asks for it)
   CU(cudaMemcpyAsync(h->h_ctrl, h->p.ctrl, offsetof(Ctrl, round_nnz), cudaMemcpyDeviceToHost, h->stream));
   h->histfetched = false;
   h->pending = true;
   return GPULIN_OK;
}

extern "C" int gpulin_propagate_wait(gpulin_t* h, gpulin_result* res)
{
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   if( !h->pending )
      return fail(GPULIN_ERR_STATE, "gpulin_propagate_wait without gpulin_propagate_async");
   CU(cudaSetDevice(h->device));
   CU(cudaStreamSynchronize(h->stream));
   h->pending = false;
   float ms = 0.0f;
   CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
   h->lastsmall = h->smallcall;
   h->lastresumed = false;
   if( h->smallcall && h->h_ctrl->status == GPULIN_PROBE_OVERFLOW )
   {
      h->lastresumed = true;
      h->smallrounds = h->h_ctrl->round;
      // the cascade outgrew the block: the general loop continues the call (begin_kernel sees Ctrl::resume)
      h->smallcall = false;
      CU(cudaEventRecord(h->ev0, h->stream));
      CU(cudaGraphLaunch(h->gexec, h->stream));       // (probe_kernel left maxrounds and logcap in the control block)
      CU(cudaEventRecord(h->ev1, h->stream));
      CU(cudaMemcpyAsync(h->h_ctrl, h->p.ctrl, offsetof(Ctrl, round_nnz), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      float ms2 = 0.0f;
      CU(cudaEventElapsedTime(&ms2, h->ev0, h->ev1));
      ms += ms2;
   }
   const Ctrl* c = h->h_ctrl;
   h->last.status = c->status;
   h->last.nrounds = c->round;
   h->last.nchanges = (int64_t)c->total_nchg;
   h->last.nnz_processed = (int64_t)c->total_nnz;
   h->last.device_ms = (double)ms;
   h->lastrounds = c->round;
   h->smallcols = (c->status == GPULIN_FIXPOINT && !c->peererror) ? 0 : -1;     // nothing is marked after a fixpoint
   if( res != nullptr )
      *res = h->last;
   if( c->peererror )
      return fail(GPULIN_ERR_STATE, "a peer rank did not deliver its candidates within 5 s");
   return GPULIN_OK;
}

extern "C" int gpulin_propagate(gpulin_t* h, int maxrounds, gpulin_result* res)
{
   OK(gpulin_propagate_async(h, maxrounds));
   return gpulin_propagate_wait(h, res);
}

// probing: start from the (propagated) state of `base`: bounds and keys are copied, nothing is marked
extern "C" int gpulin_reset_from(gpulin_t* h, gpulin_t* base)
{
   if( h != nullptr )
      ++h->version;
   if( h == nullptr || base == nullptr || h->shared != base->shared )
      return fail(GPULIN_ERR_ARG, "gpulin_reset_from needs a handle and a clone of it");
   if( !base->havebounds )
      return fail(GPULIN_ERR_STATE, "the base handle has no bounds");
   CU(cudaSetDevice(h->device));
   CU(cudaMemcpyAsync(const_cast<double2*>(h->p.bnd), base->p.bnd, sizeof(double2) * (size_t)h->ncols, cudaMemcpyDeviceToDevice, h->stream));
   CU(cudaMemcpyAsync(h->p.cand, base->p.cand, sizeof(long long) * 2 * (size_t)h->ncols, cudaMemcpyDeviceToDevice, h->stream));
   CU(cudaMemcpyAsync(h->p.freebits, base->p.freebits, h->freebytes, cudaMemcpyDeviceToDevice, h->stream));
   CU(cudaMemcpyAsync(h->p.bndf, base->p.bndf, sizeof(double2) * (size_t)h->ncols, cudaMemcpyDeviceToDevice, h->stream));
   CU(cudaMemcpyAsync(h->p.colstate, base->p.colstate, (size_t)h->ncols, cudaMemcpyDeviceToDevice, h->stream));
   CU(cudaMemcpyAsync(h->p.fracflag, base->p.fracflag, sizeof(unsigned), cudaMemcpyDeviceToDevice, h->stream));
   CU(cudaMemsetAsync(h->p.dirty, 0, (size_t)h->nrows, h->stream));
   CU(cudaMemsetAsync(&h->p.ctrl->poisoned, 0, sizeof(unsigned), h->stream));
   CU(cudaMemsetAsync(&h->p.ctrl->nmark[0][0], 0, sizeof(unsigned) * 8, h->stream));
   h->smallcols = 0;               // the state of a node whose propagation is complete
   CU(cudaMemsetAsync(h->p.tileflag, 0, (size_t)h->ntiles, h->stream));
   CU(cudaMemsetAsync(h->p.colbits, 0, sizeof(unsigned) * ((size_t)h->ncols / 32 + 1), h->stream));
   h->havebounds = true;
   return GPULIN_OK;
}

// BASELINE config 5 / SCIPapplyProbingVar (prop_probing.c:1254-1279): probe i starts from the bounds of `base` (the
// node), sets variable var[i] to [lb[i], ub[i]] and propagates to its fixpoint.  nworkers clones of the base handle keep
// that many probes in flight on their own streams; a worker returns to the node by undoing its change log.
static int probeBatchImpl(gpulin_t* base, int nworkers, int64_t nprobes, const int32_t* var, const double* lb,
   const double* ub, int maxrounds, int32_t* status, int32_t* nrounds, int64_t* nchanges, bool wantlog, int64_t* chgbeg,
   gpulin_change* chg, int64_t maxchg, int64_t* nchg)
{
   if( base == nullptr || nworkers < 1 || nprobes < 0 || (nprobes > 0 && (var == nullptr || lb == nullptr || ub == nullptr)) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( !base->havebounds )
      return fail(GPULIN_ERR_STATE, "gpulin_probe_batch before gpulin_set_bounds on the base handle");
   for( int64_t i = 0; i < nprobes; ++i )
   {
      if( var[i] < 0 || var[i] >= base->ncols )
         return fail(GPULIN_ERR_ARG, "probe %lld: column index %d out of range", (long long)i, var[i]);
   }
   CU(cudaSetDevice(base->device));
   CU(cudaStreamSynchronize(base->stream));
   const int64_t logcap = 1 << 16;
   while( (int)base->workers.size() < nworkers )
   {
      gpulin* w = nullptr;
      OK(gpulin_clone(base, &w));
      int rc = gpulin_set_change_log(w, logcap);
      if( rc != GPULIN_OK )
      {
         gpulin_destroy(w);
         return rc;
      }
      w->lightfetch = true;
      base->workers.push_back(w);
   }
   for( gpulin* w : base->workers )
   {
      if( w->syncedversion != base->version )
         w->needreset = true;       // the node has changed since this worker last took its bounds over
   }

   // ---- one launch per WORKER (probe_list_kernel): worker w runs the probes w, w + nworkers, ... one after the other, every
   // ---- probe inside one block; the probes go to the device in one copy each way
   if( nprobes > base->proberescap )
   {
      cudaFree(base->d_proberes);
      cudaFree(base->d_probevar);
      cudaFree(base->d_probelb);
      cudaFree(base->d_probeub);
      base->d_proberes = nullptr;
      base->d_probevar = nullptr;
      base->d_probelb = base->d_probeub = nullptr;
      base->proberescap = 0;
      CU(cudaMalloc(&base->d_proberes, sizeof(ProbeResult) * (size_t)nprobes));
      CU(cudaMalloc((void**)&base->d_probevar, sizeof(int) * (size_t)nprobes));
      CU(cudaMalloc((void**)&base->d_probelb, sizeof(double) * (size_t)nprobes));
      CU(cudaMalloc((void**)&base->d_probeub, sizeof(double) * (size_t)nprobes));
      base->proberescap = nprobes;
   }
   ProbeResult* d_res = (ProbeResult*)base->d_proberes;
   const bool general = false;
   if( wantlog )
   {
      const int64_t want = std::max<int64_t>(maxchg, 1);
      if( want > base->probelogcap )
      {
         cudaFree(base->d_probelog);
         base->d_probelog = nullptr;
         base->probelogcap = 0;
         CU(cudaMalloc((void**)&base->d_probelog, sizeof(ChangeRec) * (size_t)want));
         base->probelogcap = want;
      }
      if( base->d_probecursor == nullptr )
         CU(cudaMalloc((void**)&base->d_probecursor, sizeof(unsigned long long)));
      CU(cudaMemset(base->d_probecursor, 0, sizeof(unsigned long long)));
   }
   if( nprobes > 0 && !general )
   {
      CU(cudaMemcpy(base->d_probevar, var, sizeof(int) * (size_t)nprobes, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(base->d_probelb, lb, sizeof(double) * (size_t)nprobes, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(base->d_probeub, ub, sizeof(double) * (size_t)nprobes, cudaMemcpyHostToDevice));
   }
   for( int wi = 0; wi < nworkers && wi < nprobes && !general; ++wi )
   {
      gpulin* w = base->workers[(size_t)wi];
      if( w->needreset )
      {
         OK(gpulin_reset_from(w, base));
         w->lastvar = -1;
         w->needreset = false;
         w->syncedversion = base->version;
      }
      probe_list_kernel<<<1, PROBE_THREADS, 0, w->stream>>>(w->p, base->p, w->lastvar, base->d_probevar, base->d_probelb,
         base->d_probeub, wi, nworkers, (int)nprobes, maxrounds, (int)w->logcap, d_res, wantlog ? base->d_probelog : nullptr,
         base->d_probecursor, wantlog ? base->probelogcap : 0);
      w->lastvar = var[wi + ((nprobes - 1 - wi) / nworkers) * nworkers];
   }
   CU(cudaGetLastError());
   for( int wi = 0; wi < nworkers; ++wi )
      CU(cudaStreamSynchronize(base->workers[(size_t)wi]->stream));
   std::vector<ProbeResult> res((size_t)nprobes);
   if( nprobes > 0 && !general )
      CU(cudaMemcpy(res.data(), d_res, sizeof(ProbeResult) * (size_t)nprobes, cudaMemcpyDeviceToHost));
   // ---- the change logs: device order -> probe order
   std::vector<ChangeRec> devlog;
   std::vector<std::vector<ChangeRec>> rerunlog;
   if( wantlog && nprobes > 0 )
   {
      unsigned long long used = 0;
      CU(cudaMemcpy(&used, base->d_probecursor, sizeof(used), cudaMemcpyDeviceToHost));
      used = std::min<unsigned long long>(used, (unsigned long long)base->probelogcap);
      devlog.resize((size_t)used);
      if( used > 0 )
         CU(cudaMemcpy(devlog.data(), base->d_probelog, sizeof(ChangeRec) * (size_t)used, cudaMemcpyDeviceToHost));
      rerunlog.resize((size_t)nprobes);
   }
   // ---- probes that outgrew the block are rerun through the general loop on worker 0
   for( int64_t i = 0; i < nprobes; ++i )
   {
      if( !general && res[(size_t)i].status != GPULIN_PROBE_OVERFLOW )
         continue;
      if( !general )
         base->workers[(size_t)(i % nworkers)]->needreset = true;
      gpulin* w = base->workers[0];
      OK(gpulin_reset_from(w, base));
      w->needreset = true;
      update_one_kernel<<<1, 32, 0, w->stream>>>(w->p, var[i], lb[i], ub[i]);
      CU(cudaGetLastError());
      gpulin_result r;
      OK(gpulin_propagate(w, maxrounds, &r));
      res[(size_t)i].status = r.status;
      res[(size_t)i].nrounds = r.nrounds;
      res[(size_t)i].nchanges = r.nchanges;
      if( wantlog )
      {
         int64_t nl = 0;
         std::vector<ChangeRec>& rl = rerunlog[(size_t)i];
         rl.resize((size_t)std::min<int64_t>(r.nchanges, w->logcap));
         static_assert(sizeof(gpulin_change) == sizeof(ChangeRec), "change record layout");
         OK(gpulin_get_changes(w, (gpulin_change*)rl.data(), (int64_t)rl.size(), &nl));
         res[(size_t)i].logoff = -2;       // in rerunlog
         res[(size_t)i].nlog = (long long)std::min<int64_t>(nl, (int64_t)rl.size());
      }
   }
   for( int64_t i = 0; i < nprobes; ++i )
   {
      if( status != nullptr ) status[i] = res[(size_t)i].status;
      if( nrounds != nullptr ) nrounds[i] = res[(size_t)i].nrounds;
      if( nchanges != nullptr ) nchanges[i] = res[(size_t)i].nchanges;
   }
   if( wantlog )
   {
      int64_t pos = 0;
      bool fits = true;
      for( int64_t i = 0; i < nprobes; ++i )
      {
         const ProbeResult& pr = res[(size_t)i];
         chgbeg[i] = pos;
         const ChangeRec* src = nullptr;
         if( pr.logoff == -2 )
            src = rerunlog[(size_t)i].data();
         else if( pr.logoff >= 0 && (size_t)(pr.logoff + pr.nlog) <= devlog.size() )
            src = devlog.data() + pr.logoff;
         else if( pr.nlog > 0 )
            fits = false;                   // the device buffer was too small for this probe
         if( pos + pr.nlog <= maxchg && src != nullptr && fits )
            memcpy((void*)(chg + pos), src, sizeof(ChangeRec) * (size_t)pr.nlog);
         else if( pr.nlog > 0 )
            fits = false;
         pos += pr.nlog;
      }
      chgbeg[nprobes] = pos;
      *nchg = pos;
      if( !fits )
         return fail(GPULIN_ERR_ARG, "the change buffer holds %lld entries, the batch produced %lld", (long long)maxchg, (long long)pos);
   }
   return GPULIN_OK;
}

extern "C" int gpulin_probe_batch(gpulin_t* base, int nworkers, int64_t nprobes, const int32_t* var, const double* lb,
   const double* ub, int maxrounds, int32_t* status, int32_t* nrounds, int64_t* nchanges)
{
   return probeBatchImpl(base, nworkers, nprobes, var, lb, ub, maxrounds, status, nrounds, nchanges, false, nullptr, nullptr, 0, nullptr);
}

extern "C" int gpulin_probe_batch_changes(gpulin_t* base, int nworkers, int64_t nprobes, const int32_t* var, const double* lb,
   const double* ub, int maxrounds, int32_t* status, int32_t* nrounds, int64_t* nchanges, int64_t* chgbeg, gpulin_change* chg,
   int64_t maxchg, int64_t* nchg)
{
   if( chgbeg == nullptr || nchg == nullptr || maxchg < 0 || (maxchg > 0 && chg == nullptr) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   return probeBatchImpl(base, nworkers, nprobes, var, lb, ub, maxrounds, status, nrounds, nchanges, true, chgbeg, chg, maxchg, nchg);
}

// a second set of bound vectors on the same matrix: shares the read-only arrays, owns everything a round writes
extern "C" int gpulin_clone(gpulin_t* src, gpulin_t** out)
{
   if( src == nullptr || out == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   *out = nullptr;
   CU(cudaSetDevice(src->device));
   gpulin* h = new gpulin();
   h->shared = src->shared;
   ++h->shared->refs;
   h->device = src->device;
   h->nrows = src->nrows; h->ncols = src->ncols; h->nnz = src->nnz; h->nstored = src->nstored;
   h->nsell = src->nsell; h->nstream = src->nstream; h->nlong = src->nlong; h->ntiles = src->ntiles;
   h->nstreamelems = src->nstreamelems; h->maxlen = src->maxlen;
   h->nsellblocks = src->nsellblocks; h->nstreamblocks = src->nstreamblocks;
   h->nlongblocks = src->nlongblocks; h->napplyblocks = src->napplyblocks; h->nexactblocks = src->nexactblocks; h->exactkernel = src->exactkernel;
   h->nsparseblocks = src->nsparseblocks;
   h->npushblocks = src->npushblocks;
   h->nsm = src->nsm; h->hostloop = src->hostloop; h->perm = src->perm;
   h->p = src->p;
   DevProblem& p = h->p;
   unsigned char* d_dirty; unsigned char* d_tileflag; int* d_xlist; double2* d_bnd; long long* d_cand; unsigned* d_colbits;
   int* d_chglist; Ctrl* d_ctrl; int* d_marklist; unsigned* d_freebits; double2* d_bndf;
   unsigned char* d_colstate; int* d_flist; FastAcc* d_facc; unsigned* d_fracflag;
   int rc = GPULIN_OK;
   TRY(devAlloc(h, &d_colstate, (size_t)h->ncols + 1));
   TRY(devAlloc(h, &d_flist, (size_t)h->nsell + 1));
   TRY(devAlloc(h, &d_facc, (size_t)h->nsell + 1));
   TRY(devAlloc(h, &d_fracflag, 1));
   TRYCU(cudaMemset(d_colstate, CS_OTHER, (size_t)h->ncols + 1));
   TRYCU(cudaMemset(d_fracflag, 0, sizeof(unsigned)));
   TRY(devAlloc(h, &d_dirty, (size_t)h->nrows + 64));
   TRY(devAlloc(h, &d_tileflag, (size_t)h->ntiles + 64));
   TRY(devAlloc(h, &d_xlist, (size_t)h->nrows + 1));
   TRY(devAlloc(h, &d_marklist, (size_t)6 * MARKCAP));
   TRY(devAlloc(h, &d_bnd, (size_t)h->ncols + 1));
   h->freebytes = src->freebytes;
   h->nsellbitsblocks = src->nsellbitsblocks;
   h->sellbitsvariant = src->sellbitsvariant;
   TRY(devAlloc(h, &d_freebits, h->freebytes / sizeof(unsigned)));
   TRY(devAlloc(h, &d_bndf, (size_t)h->ncols + 1));
   TRY(devAlloc(h, &d_cand, 2 * (size_t)h->ncols + 2));
   TRY(devAlloc(h, &d_colbits, (size_t)h->ncols / 32 + 2));
   TRY(devAlloc(h, &d_chglist, (size_t)h->ncols + 1));
   TRY(devAlloc(h, &d_ctrl, 1));
   TRY(devAlloc(h, &h->d_peers, 1));
   TRY(devAlloc(h, &h->d_tmplb, (size_t)h->ncols + 1));
   TRY(devAlloc(h, &h->d_tmpub, (size_t)h->ncols + 1));
   TRYCU(cudaMemset(d_dirty, 0, (size_t)h->nrows + 64));
   TRYCU(cudaMemset(d_tileflag, 0, (size_t)h->ntiles + 64));
   TRYCU(cudaMemset(d_colbits, 0, sizeof(unsigned) * ((size_t)h->ncols / 32 + 2)));
   TRYCU(cudaMemset(d_ctrl, 0, sizeof(Ctrl)));
   TRYCU(cudaMemset(d_cand, 0, sizeof(long long) * (2 * (size_t)h->ncols + 2)));
   TRYCU(cudaMallocHost((void**)&h->h_ctrl, sizeof(Ctrl)));
   TRYCU(cudaMallocHost((void**)&h->h_params, 4 * sizeof(int)));
   if( rc == GPULIN_OK )
      memset(h->h_ctrl, 0, sizeof(Ctrl));
   TRYCU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
   TRYCU(cudaStreamCreateWithFlags(&h->aux[0], cudaStreamNonBlocking));
   TRYCU(cudaStreamCreateWithFlags(&h->aux[1], cudaStreamNonBlocking));
   TRYCU(cudaEventCreateWithFlags(&h->evfork, cudaEventDisableTiming));
   TRYCU(cudaEventCreateWithFlags(&h->evjoin[0], cudaEventDisableTiming));
   TRYCU(cudaEventCreateWithFlags(&h->evjoin[1], cudaEventDisableTiming));
   TRYCU(cudaEventCreateWithFlags(&h->evfork2, cudaEventDisableTiming));
   TRYCU(cudaEventCreateWithFlags(&h->evjoin2, cudaEventDisableTiming));
   TRYCU(cudaEventCreate(&h->ev0));
   TRYCU(cudaEventCreate(&h->ev1));
   if( rc != GPULIN_OK )
   {
      gpulin_destroy(h);
      return rc;
   }
   p.dirty = d_dirty;
   p.tileflag = d_tileflag;
   p.xlist = d_xlist;
   p.marklist = d_marklist;
   p.bnd = d_bnd;
   p.freebits = d_freebits;
   p.bndf = d_bndf;
   p.colstate = d_colstate;
   p.flist = d_flist;
   p.facc = d_facc;
   p.fracflag = d_fracflag;
   p.cand = d_cand;
   p.colbits = d_colbits;
   p.chglist = d_chglist;
   p.ctrl = d_ctrl;
   p.log = nullptr;
   p.peers = nullptr;
   h->nrangedblocks = src->nrangedblocks;
   if( p.rr.n > 0 )
   {
      // the ranged rows are shared, the scratch words of the rule are not
      const size_t words = (size_t)h->nsm * 64 * (size_t)p.rr.scratchwords;
      if( cudaMalloc(&h->d_rr[5], sizeof(unsigned) * words) != cudaSuccess )
      {
         gpulin_destroy(h);
         return fail(GPULIN_ERR_NOMEM, "cudaMalloc of the ranged-row scratch failed");
      }
      p.rr.scratch = (unsigned*)h->d_rr[5];
   }
   // (the clone of a handle that shares its dense rounds with peers works alone: it takes every row itself)
   h->npeers = 1;
   h->peerrank = 0;
   setShare(h, 0, 1);
   if( !h->hostloop )
   {
      rc = buildGraph(h);
      if( rc != GPULIN_OK )
      {
         gpulin_destroy(h);
         return rc;
      }
   }
   *out = h;
   return GPULIN_OK;
}

extern "C" int gpulin_get_bounds_device(gpulin_t* h, double* d_lb, double* d_ub)
{
   if( h == nullptr || d_lb == nullptr || d_ub == nullptr )
      return fail(GPULIN_ERR_ARG, "NULL argument");
   CU(cudaSetDevice(h->device));
   get_bounds_kernel<<<gridFor(h, h->ncols), 256, 0, h->stream>>>(h->p, d_lb, d_ub);
   CU(cudaGetLastError());
   return GPULIN_OK;
}

extern "C" int gpulin_get_bounds(gpulin_t* h, double* lb, double* ub)
{
   if( h == nullptr || lb == nullptr || ub == nullptr )
      return fail(GPULIN_ERR_ARG, "NULL argument");
   OK(gpulin_get_bounds_device(h, h->d_tmplb, h->d_tmpub));
   CU(cudaMemcpyAsync(lb, h->d_tmplb, sizeof(double) * (size_t)h->ncols, cudaMemcpyDeviceToHost, h->stream));
   CU(cudaMemcpyAsync(ub, h->d_tmpub, sizeof(double) * (size_t)h->ncols, cudaMemcpyDeviceToHost, h->stream));
   CU(cudaStreamSynchronize(h->stream));
   return GPULIN_OK;
}

extern "C" int gpulin_set_change_log(gpulin_t* h, int64_t capacity)
{
   if( h == nullptr || capacity < 0 || capacity >= (1LL << 31) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   CU(cudaSetDevice(h->device));
   CU(cudaStreamSynchronize(h->stream));
   if( h->d_log != nullptr )
   {
      cudaFree(h->d_log);
      h->d_log = nullptr;
   }
   h->logcap = 0;
   h->p.log = nullptr;
   if( capacity > 0 )
   {
      cudaError_t e = cudaMalloc((void**)&h->d_log, sizeof(ChangeRec) * (size_t)capacity);
      if( e != cudaSuccess )
         return fail(GPULIN_ERR_NOMEM, "cudaMalloc of the change log failed: %s", cudaGetErrorString(e));
      h->logcap = capacity;
      h->p.log = h->d_log;
   }
   // kernel parameters are baked into the graph
   if( !h->hostloop )
      OK(buildGraph(h));
   return GPULIN_OK;
}

extern "C" int gpulin_get_changes(gpulin_t* h, gpulin_change* out, int64_t maxn, int64_t* n)
{
   if( h == nullptr || n == nullptr || maxn < 0 || (maxn > 0 && out == nullptr) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   CU(cudaSetDevice(h->device));
   const int64_t produced = (int64_t)h->h_ctrl->logcount;
   *n = produced;
   const int64_t m = std::min(std::min(produced, h->logcap), maxn);
   if( m > 0 )
   {
      static_assert(sizeof(gpulin_change) == sizeof(ChangeRec), "change record layout");
      CU(cudaMemcpyAsync(out, h->d_log, sizeof(ChangeRec) * (size_t)m, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
   }
   return GPULIN_OK;
}

// rows that are redundant for the bounds on the device, in the caller's row numbering, ascending
extern "C" int gpulin_get_redundant_rows(gpulin_t* h, int32_t* rows, int64_t maxn, int64_t* n)
{
   if( h == nullptr || n == nullptr || maxn < 0 || (maxn > 0 && rows == nullptr) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( !h->havebounds )
      return fail(GPULIN_ERR_STATE, "no bounds on the device");
   *n = 0;
   if( h->nrows == 0 )
      return GPULIN_OK;
   CU(cudaSetDevice(h->device));
   if( h->d_redflags == nullptr )
      CU(cudaMalloc((void**)&h->d_redflags, (size_t)h->nrows));
   if( h->h_redflags == nullptr )
      CU(cudaMallocHost((void**)&h->h_redflags, (size_t)h->nrows));
   const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((h->nrows + 7) / 8, (int64_t)h->nsm * 8));
   redundant_rows_kernel<<<blocks, 256, 0, h->stream>>>(h->p, h->d_redflags);
   CU(cudaMemcpyAsync(h->h_redflags, h->d_redflags, (size_t)h->nrows, cudaMemcpyDeviceToHost, h->stream));
   CU(cudaStreamSynchronize(h->stream));
   std::vector<int32_t> found;
   for( int64_t i = 0; i < h->nrows; ++i )
   {
      if( h->h_redflags[(size_t)i] )
         found.push_back(h->perm[(size_t)i]);
   }
   std::sort(found.begin(), found.end());
   *n = (int64_t)found.size();
   for( int64_t i = 0; i < *n && i < maxn; ++i )
      rows[i] = found[(size_t)i];
   return GPULIN_OK;
}

extern "C" int gpulin_get_round_stats(gpulin_t* h, double* ms, int64_t* nnz, int64_t* nchg, int32_t maxn, int32_t* n)
{
   if( h == nullptr || n == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( !h->histfetched )
   {
      CU(cudaSetDevice(h->device));
      CU(cudaMemcpyAsync(h->h_ctrl, h->p.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      h->histfetched = true;
   }
   const Ctrl* c = h->h_ctrl;
   const int m = std::min(std::min(h->lastrounds, (int)MAX_HIST), (int)maxn);
   unsigned long long prev = c->t_start;
   for( int i = 0; i < m; ++i )
   {
      if( ms != nullptr )
         ms[i] = 1e-6 * (double)(c->hist_time[i] - prev);
      if( nnz != nullptr )
         nnz[i] = (int64_t)c->hist_nnz[i];
      if( nchg != nullptr )
         nchg[i] = (int64_t)c->hist_nchg[i];
      prev = c->hist_time[i];
   }
   *n = m;
   return GPULIN_OK;
}

extern "C" int gpulin_get_exchange_stats(gpulin_t* h, double* before_ms, double* wait_ms, int32_t maxn, int32_t* n)
{
   if( h == nullptr || n == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( !h->histfetched )
   {
      CU(cudaSetDevice(h->device));
      CU(cudaMemcpyAsync(h->h_ctrl, h->p.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      h->histfetched = true;
   }
   const Ctrl* c = h->h_ctrl;
   const int m = std::min(std::min(h->lastrounds, (int)MAX_HIST), (int)maxn);
   unsigned long long prev = c->t_start;
   for( int i = 0; i < m; ++i )
   {
      const bool has = h->npeers > 1 && c->hist_push[i] != 0;
      if( before_ms != nullptr )
         before_ms[i] = has ? 1e-6 * (double)(c->hist_push[i] - prev) : -1.0;
      if( wait_ms != nullptr )
         wait_ms[i] = has ? 1e-6 * (double)c->hist_wait[i] : 0.0;
      prev = c->hist_time[i];
   }
   *n = m;
   return GPULIN_OK;
}

extern "C" int gpulin_get_trace(gpulin_t* h, int32_t* ids, double* us, int32_t maxn, int32_t* n)
{
   if( h == nullptr || n == nullptr || (maxn > 0 && (ids == nullptr || us == nullptr)) )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   if( !h->histfetched )
   {
      CU(cudaSetDevice(h->device));
      CU(cudaMemcpyAsync(h->h_ctrl, h->p.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      h->histfetched = true;
   }
   const Ctrl* c = h->h_ctrl;
   const int m = std::min(std::min((int)c->ntrace, (int)MAX_TRACE), (int)maxn);
   const unsigned long long mask = 0x00ffffffffffffffull;
   // (the stamps are appended by atomics in launch order, but kernels on side streams may interleave: sort by time)
   std::vector<unsigned long long> ev(c->trace, c->trace + m);
   std::sort(ev.begin(), ev.end(), [&](unsigned long long a, unsigned long long b) { return (a & mask) < (b & mask); });
   for( int i = 0; i < m; ++i )
   {
      ids[i] = (int32_t)(ev[(size_t)i] >> 56);
      us[i] = 1e-3 * (double)((ev[(size_t)i] & mask) - (c->t_start & mask));
   }
   *n = m;
   return GPULIN_OK;
}

extern "C" int gpulin_get_layout(gpulin_t* h, int64_t* stats, int32_t nstats)
{
   if( h == nullptr || stats == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   const int64_t v[12] = {h->nnz, h->nstored, h->nsell, h->nstream, h->nlong, (int64_t)h->devbytes, h->ntiles,
      h->nsellblocks, h->nstreamblocks, h->maxlen, h->nsellunit, h->nsellbitsblocks};
   for( int i = 0; i < nstats && i < 12; ++i )
      stats[i] = v[i];
   return GPULIN_OK;
}

extern "C" int gpulin_get_call_stats(gpulin_t* h, int64_t* stats, int32_t nstats)
{
   if( h == nullptr || stats == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   const Ctrl* c = h->h_ctrl;
   const int64_t rounds = h->lastrounds;
   const int nkinds = (h->nsellblocks > 0) + (h->nstreamblocks > 0) + (h->nlongblocks > 0);
   int64_t launches = 0;
   int64_t dense = 0;
   int64_t sparse = 0;
   if( h->lastsmall && !h->lastresumed )
      launches = 1;                                      // probe_kernel ran every round
   else
   {
      sparse = c->nsparse;
      dense = rounds - sparse - (h->lastresumed ? (int64_t)h->smallrounds : 0);
      if( dense < 0 )
         dense = 0;
      // begin; per dense round: sweeps, exact, collect, [merge with peers], apply, sparse rounds
      // (the graph runs its first round unconditionally; fast_rows_kernel is part of that round only)
      const int64_t looprounds = std::max<int64_t>(dense, 1);
      launches = (h->lastresumed ? 1 : 0) + 1 + (h->nfastblocks > 0 ? 1 : 0)
         + looprounds * (nkinds + 3 + (h->npeers > 1 ? 1 : 0) + (h->nsparseblocks > 0 ? 1 : 0));
   }
   const int64_t v[6] = {launches, dense, sparse, h->lastsmall ? 1 : 0, h->lastresumed ? 1 : 0, (int64_t)c->nfastrows};
   for( int i = 0; i < nstats && i < 6; ++i )
      stats[i] = v[i];
   return GPULIN_OK;
}

extern "C" int gpulin_algorithmic_bytes(gpulin_t* h, int64_t* bytes)
{
   if( h == nullptr || bytes == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   *bytes = h->nnz * 12 + h->nrows * 20 + h->ncols * 17;
   return GPULIN_OK;
}

// ---- single steps (multi-GPU rounds, profiling) -----------------------------------------------------------------

extern "C" int gpulin_exchange_buffer(gpulin_t* h, int64_t** d_keys, int64_t* nkeys)
{
   if( h == nullptr || d_keys == nullptr || nkeys == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   *d_keys = (int64_t*)h->p.cand;
   *nkeys = 2 * h->ncols + 2;
   return GPULIN_OK;
}

extern "C" int gpulin_get_keys(gpulin_t* h, int64_t* keys)
{
   if( h == nullptr || keys == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   CU(cudaSetDevice(h->device));
   CU(cudaMemcpyAsync(keys, h->p.cand, sizeof(long long) * (2 * (size_t)h->ncols + 2), cudaMemcpyDeviceToHost, h->stream));
   CU(cudaStreamSynchronize(h->stream));
   return GPULIN_OK;
}

extern "C" int gpulin_set_keys(gpulin_t* h, const int64_t* keys)
{
   if( h != nullptr )
      ++h->version;
   if( h == nullptr || keys == nullptr )
      return fail(GPULIN_ERR_ARG, "invalid argument");
   CU(cudaSetDevice(h->device));
   CU(cudaMemcpyAsync(h->p.cand, keys, sizeof(long long) * (2 * (size_t)h->ncols + 2), cudaMemcpyHostToDevice, h->stream));
   CU(cudaStreamSynchronize(h->stream));
   return GPULIN_OK;
}

extern "C" int gpulin_mark_all(gpulin_t* h)
{
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   CU(cudaSetDevice(h->device));
   mark_all_kernel<<<gridFor(h, h->nrows), 256, 0, h->stream>>>(h->p);
   h->smallcols = -1;
   CU(cudaGetLastError());
   return GPULIN_OK;
}

extern "C" int gpulin_round_begin(gpulin_t* h)
{
   if( h != nullptr )
      h->smallcols = -1;
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   if( !h->havebounds )
      return fail(GPULIN_ERR_STATE, "gpulin_round_begin before gpulin_set_bounds");
   CU(cudaSetDevice(h->device));
   h->h_params[0] = 0;
   h->h_params[1] = (int)h->logcap;
   CU(cudaMemcpyAsync(&h->p.ctrl->maxrounds, h->h_params, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
   begin_kernel<<<1, 1, 0, h->stream>>>(h->p.ctrl);
   CU(cudaGetLastError());
   return GPULIN_OK;
}

extern "C" int gpulin_round_sweep(gpulin_t* h)
{
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   if( !h->havebounds )
      return fail(GPULIN_ERR_STATE, "gpulin_round_sweep before gpulin_set_bounds");
   CU(cudaSetDevice(h->device));
   OK((launchRoundKernels<APPLY_LIST, false>(h, true, false)));
   publish_cutoff_kernel<<<1, 1, 0, h->stream>>>(h->p);
   CU(cudaGetLastError());
   return GPULIN_OK;
}

extern "C" int gpulin_round_apply(gpulin_t* h, int dense, int64_t* nchanges, int32_t* cutoff)
{
   if( h != nullptr )
      ++h->version;
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   CU(cudaSetDevice(h->device));
   absorb_cutoff_kernel<<<1, 1, 0, h->stream>>>(h->p);
   if( dense )
      OK((launchRoundKernels<APPLY_DENSE, false>(h, false, true)));
   else
      OK((launchRoundKernels<APPLY_LIST, false>(h, false, true)));
   if( nchanges != nullptr || cutoff != nullptr )
   {
      OK(fetchCtrl(h));
      const Ctrl* c = h->h_ctrl;
      const int r = c->round - 1;
      if( nchanges != nullptr )
         *nchanges = (r >= 0 && r < MAX_HIST) ? (int64_t)c->hist_nchg[r] : -1;
      if( cutoff != nullptr )
         *cutoff = c->cutoff;
      h->lastrounds = c->round;
      h->last.status = c->status;
      h->last.nrounds = c->round;
      h->last.nchanges = (int64_t)c->total_nchg;
      h->last.nnz_processed = (int64_t)c->total_nnz;
   }
   return GPULIN_OK;
}

// one full round (every row marked) with CUDA events between its kernels: filter sweep(s) | exact kernel | apply
extern "C" int gpulin_profile_round(gpulin_t* h, double* sweep_ms, double* exact_ms, double* apply_ms)
{
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   if( !h->havebounds )
      return fail(GPULIN_ERR_STATE, "gpulin_profile_round before gpulin_set_bounds");
   CU(cudaSetDevice(h->device));
   if( h->evprof[0] == nullptr )
   {
      for( int i = 0; i < 4; ++i )
         CU(cudaEventCreate(&h->evprof[i]));
   }
   OK(gpulin_round_begin(h));
   OK(gpulin_mark_all(h));
   CU(cudaEventRecord(h->evprof[0], h->stream));
   const int keepexact = h->nexactblocks;
   h->nexactblocks = 0;                               // launchRoundKernels skips the exact kernel ...
   OK((launchRoundKernels<APPLY_LIST, false>(h, true, false, false)));
   h->nexactblocks = keepexact;
   CU(cudaEventRecord(h->evprof[1], h->stream));
   OK(launchExact(h, false));                         // ... which is timed on its own (as in the loop body)
   CU(cudaEventRecord(h->evprof[2], h->stream));
   OK(launchCollect(h));                              // (counted with the apply stage)
   OK((launchRoundKernels<APPLY_LIST, false>(h, false, true)));
   CU(cudaEventRecord(h->evprof[3], h->stream));
   CU(cudaEventSynchronize(h->evprof[3]));
   float t[3] = {0.f, 0.f, 0.f};
   for( int i = 0; i < 3; ++i )
      CU(cudaEventElapsedTime(&t[i], h->evprof[i], h->evprof[i + 1]));
   if( sweep_ms != nullptr ) *sweep_ms = t[0];
   if( exact_ms != nullptr ) *exact_ms = t[1];
   if( apply_ms != nullptr ) *apply_ms = t[2];
   OK(fetchCtrl(h));
   return GPULIN_OK;
}

// ---- dense rounds shared by the GPUs of one node (the matrix and the bounds are replicated) -------------------------------

static int allocInbox(gpulin* h, int nranks)
{
   if( h->d_inbox != nullptr && h->inboxranks >= nranks )
      return GPULIN_OK;
   if( h->npeers > 1 )
      return fail(GPULIN_ERR_STATE, "the handle is already connected");
   CU(cudaSetDevice(h->device));
   cudaFree(h->d_inbox);
   h->d_inbox = nullptr;
   const long long cap = (h->ncols + 3) & ~3LL;
   const size_t bytes = peerBoxBytes(nranks, cap);
   cudaError_t e = cudaMalloc((void**)&h->d_inbox, bytes);
   if( e != cudaSuccess )
      return fail(GPULIN_ERR_NOMEM, "cudaMalloc of the peer inbox (%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
   CU(cudaMemset(h->d_inbox, 0, PEER_HDR_BYTES));
   h->inboxranks = nranks;
   h->devbytes += bytes;
   return GPULIN_OK;
}

// boxes[r]: the inbox of rank r as this device addresses it
static int wirePeers(gpulin* h, int rank, int nranks, unsigned char* const* boxes)
{
   PeerTable t;
   memset(&t, 0, sizeof(t));
   t.n = nranks;
   t.rank = rank;
   t.cap = (h->ncols + 3) & ~3LL;
   for( int r = 0; r < nranks; ++r )
      t.box[r] = boxes[r];
   CU(cudaSetDevice(h->device));
   CU(cudaStreamSynchronize(h->stream));
   CU(cudaMemcpy(h->d_peers, &t, sizeof(t), cudaMemcpyHostToDevice));
   h->p.peers = h->d_peers;
   h->npeers = nranks;
   h->peerrank = rank;
   setShare(h, rank, nranks);
   h->npushblocks = (int)std::max<int64_t>(1, std::min<int64_t>((std::max(h->ncols / 32, h->nrows / 16) + 255) / 256, (int64_t)h->nsm * 4));
   // the grids of the sweeps follow the share
   {
      const int wpb = SWEEP_THREADS / 32;
      const int64_t mytiles = h->p.st1 - h->p.st0;
      const int64_t need = (mytiles + 2 * wpb - 1) / (2 * wpb);
      h->nstreamblocks = (int)std::min<int64_t>(need, (int64_t)h->nstreamblocks);
      const int64_t nlongmine = (h->nlong - rank + nranks - 1) / nranks;
      h->nlongblocks = (int)std::min<int64_t>(std::max<int64_t>(nlongmine, h->nlong > 0 ? 1 : 0), (int64_t)h->nlongblocks);
   }
   CU(cudaFuncSetAttribute(collect_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, COLLECT_SMEM_PEERS));
   if( !h->hostloop )
      OK(buildGraph(h));
   return GPULIN_OK;
}

extern "C" int gpulin_peer_handles(gpulin_t* h, int nranks, void* out, int64_t* nbytes)
{
   if( h == nullptr || nbytes == nullptr || nranks < 1 || nranks > MAX_PEERS )
      return fail(GPULIN_ERR_ARG, "invalid argument (at most %d ranks)", MAX_PEERS);
   *nbytes = (int64_t)sizeof(cudaIpcMemHandle_t);
   if( out == nullptr )
      return GPULIN_OK;
   OK(allocInbox(h, nranks));
   CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out, h->d_inbox));
   return GPULIN_OK;
}

extern "C" int gpulin_peer_connect(gpulin_t* h, int rank, int nranks, const void* allhandles)
{
   if( h == nullptr || allhandles == nullptr || nranks < 1 || nranks > MAX_PEERS || rank < 0 || rank >= nranks )
      return fail(GPULIN_ERR_ARG, "invalid argument (at most %d ranks)", MAX_PEERS);
   if( h->npeers > 1 )
      return fail(GPULIN_ERR_STATE, "gpulin_peer_connect called twice");
   if( h->d_inbox == nullptr || h->inboxranks < nranks )
      return fail(GPULIN_ERR_STATE, "gpulin_peer_connect before gpulin_peer_handles for %d ranks", nranks);
   if( nranks == 1 )
      return GPULIN_OK;
   CU(cudaSetDevice(h->device));
   const cudaIpcMemHandle_t* hd = (const cudaIpcMemHandle_t*)allhandles;
   unsigned char* boxes[MAX_PEERS] = {nullptr};
   for( int r = 0; r < nranks; ++r )
   {
      if( r == rank )
      {
         boxes[r] = h->d_inbox;
         continue;
      }
      cudaError_t e = cudaIpcOpenMemHandle(&h->peerptr[r], hd[r], cudaIpcMemLazyEnablePeerAccess);
      if( e != cudaSuccess )
         return fail(GPULIN_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s (peer access between the GPUs is required)",
            r, cudaGetErrorString(e));
      boxes[r] = (unsigned char*)h->peerptr[r];
   }
   return wirePeers(h, rank, nranks, boxes);
}

// one process, n handles of the same problem on n different devices (the SCIP plugin: SCIP is one process)
extern "C" int gpulin_group_connect(gpulin_t** hs, int n)
{
   if( hs == nullptr || n < 1 || n > MAX_PEERS )
      return fail(GPULIN_ERR_ARG, "invalid argument (at most %d handles)", MAX_PEERS);
   for( int i = 0; i < n; ++i )
   {
      if( hs[i] == nullptr || hs[i]->npeers > 1 )
         return fail(GPULIN_ERR_ARG, "handle %d is NULL or already connected", i);
      if( hs[i]->nrows != hs[0]->nrows || hs[i]->ncols != hs[0]->ncols || hs[i]->nnz != hs[0]->nnz )
         return fail(GPULIN_ERR_ARG, "handle %d holds a different problem", i);
      for( int k = 0; k < i; ++k )
      {
         if( hs[k]->device == hs[i]->device )
            return fail(GPULIN_ERR_ARG, "handles %d and %d are on the same device", k, i);
      }
   }
   if( n == 1 )
      return GPULIN_OK;
   unsigned char* boxes[MAX_PEERS] = {nullptr};
   for( int i = 0; i < n; ++i )
   {
      OK(allocInbox(hs[i], n));
      boxes[i] = hs[i]->d_inbox;
   }
   for( int i = 0; i < n; ++i )
   {
      CU(cudaSetDevice(hs[i]->device));
      for( int k = 0; k < n; ++k )
      {
         if( k == i )
            continue;
         int can = 0;
         CU(cudaDeviceCanAccessPeer(&can, hs[i]->device, hs[k]->device));
         if( !can )
            return fail(GPULIN_ERR_CUDA, "device %d cannot access the memory of device %d", hs[i]->device, hs[k]->device);
         cudaError_t e = cudaDeviceEnablePeerAccess(hs[k]->device, 0);
         if( e == cudaErrorPeerAccessAlreadyEnabled )
            (void)cudaGetLastError();
         else if( e != cudaSuccess )
            return fail(GPULIN_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", hs[i]->device, hs[k]->device, cudaGetErrorString(e));
      }
   }
   for( int i = 0; i < n; ++i )
      OK(wirePeers(hs[i], i, n, boxes));
   return GPULIN_OK;
}

extern "C" int gpulin_sync(gpulin_t* h)
{
   if( h == nullptr )
      return fail(GPULIN_ERR_ARG, "handle is NULL");
   CU(cudaSetDevice(h->device));
   CU(cudaStreamSynchronize(h->stream));
   return GPULIN_OK;
}
