// Execution plan of the WGAN-GP training hot path: parameter/workspace layout and the kernel
// sequences of one discriminator step and one generator step (train:362-368), built from the
// kernels in gemm.cu / attn.cu / lstm.cu / misc.cu.  The maths follows tests/plan_mirror.py
// (hoisted attention projection, pass batching, reverse over a forward tangent for the
// gradient penalty); every sequence is enqueued on the caller's stream, nothing allocates.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "../../include/sgg_b200.h"

namespace sgg {

static inline long long rup(long long x, long long m) { return (x + m - 1) / m * m; }

// Set while sgg_train_iteration enqueues its kernels: the annotation tensors are inputs of the whole iteration (no
// kernel of the iteration writes them), so the attention kernels may start fetching their tiles before the
// preceding kernel of the stream has finished (see pdl_wait in attn.cu).  The op-level and step-level entry points
// leave it off: their caller may have produced the annotations with the immediately preceding kernel.
static thread_local bool t_ann_static = false;
struct AnnStaticScope {
  bool prev;
  AnnStaticScope() : prev(t_ann_static) { t_ann_static = true; }
  ~AnnStaticScope() { t_ann_static = prev; }
};

// ---------------------------------------------------------------------------- row-sharded attention projection
// Data parallelism keeps every tensor replicated EXCEPT the annotation rows W_a of the attention kernel (85 % of the
// parameters): with world > 1 rank r owns the contraction indices [r*Ks, (r+1)*Ks) of flat(a) W_a for the GLOBAL batch.
// Once per iteration the ranks exchange annotation column slabs (all-to-all over NVLink); per step the only traffic of
// this block is a reduce-scatter of the partial projections P [B_global, R] and an all-gather of P_bar -- instead of
// an all-reduce of the 79 MB gradient -- and W_a's HBM traffic (K1 operand, dW_a, Adam) shrinks by 1/world per GPU.
struct ShardCtx {
  void* comm; int rank, world;
  long long Ks;                                 // contraction indices per rank (multiple of 64)
  const __nv_bfloat16* slab_g; const __nv_bfloat16* slab_d;   // [B*world, Ks]
  __nv_bfloat16* send;                          // [world][B][Ks] all-to-all staging
  // per network (index 0 = generator, 1 = discriminator): the two networks' projections run on different streams
  float* Ppart[2]; float* PBall[2]; __nv_bfloat16* PBHall[2];   // [B*world, RP] / [B*world, 2RP]
};
static thread_local const ShardCtx* t_shard = nullptr;
struct ShardScope {
  const ShardCtx* prev;
  explicit ShardScope(const ShardCtx* c) : prev(t_shard) { t_shard = c; }
  ~ShardScope() { t_shard = prev; }
};
int comm_allreduce(void* comm, float* buf, long long n, cudaStream_t st, int lane);
int comm_reduce_scatter(void* comm, const float* in, float* out, long long n_per_rank, cudaStream_t st, int lane);
int comm_all_gather(void* comm, const float* in, float* out, long long n_per_rank, cudaStream_t st, int lane);
int comm_all_to_all(void* comm, const void* send, void* recv, long long bytes_per_rank, cudaStream_t st, int lane);
bool comm_has_side_lane(void* comm);
int comm_rank(void* comm);
int comm_world(void* comm);

// ---------------------------------------------------------------------------- side stream (fork / join)
// The W_a chain of an optimiser step -- dW_a GEMM (HBM-bound, 79 MB written), its all-reduce and the Adam update of
// W_a (85 % of the parameters) -- is independent of the other weight-gradient GEMMs.  It runs on a library-owned side
// stream that forks from the caller's stream after the reverse time loop and joins it again before the next step, so
// the HBM-bound chain overlaps the tensor-bound one.  Fork and join are event edges, hence capturable into a graph.
// SGG_SIDE_STREAM=0 keeps everything on the caller's stream.
struct SideStream { cudaStream_t s; cudaEvent_t fork, join; bool ok; };
// idx 0: the W_a chain of an optimiser step; idx 1: the auxiliary stream of the two-stream reverse pass (disc_step_core)
static SideStream* side_stream(int idx = 0) {
  static thread_local SideStream ss[16][2] = {};
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("SGG_SIDE_STREAM"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& x = ss[dev][idx];
  if (!x.ok) {
    if (cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    x.ok = true;
  }
  return &x;
}
// side stream starts after everything enqueued on `st` so far
static int side_fork(cudaStream_t st, cudaStream_t* out, int idx = 0) {
  SideStream* ss = side_stream(idx);
  if (!ss) { *out = st; return 0; }
  SGG_CUDA(cudaEventRecord(ss->fork, st));
  SGG_CUDA(cudaStreamWaitEvent(ss->s, ss->fork, 0));
  *out = ss->s;
  return 0;
}
// `st` continues after everything enqueued on the side stream so far
static int side_join(cudaStream_t st, cudaStream_t s1, int idx = 0) {
  if (s1 == st) return 0;
  SideStream* ss = side_stream(idx);
  SGG_CUDA(cudaEventRecord(ss->join, s1));
  SGG_CUDA(cudaStreamWaitEvent(st, ss->join, 0));
  return 0;
}

// ============================================================================ parameter layout
struct ParamLayout {
  bool gen;
  int R, C, H, U, OUT, V, E, KX;
  long long Watt, batt, K, lng[5], lnb[5], Wdec, bdec, Wemb, total;   // fp32 bucket offsets (floats)
  // bf16 shadow: each GEMM weight is stored as [hi rows | lo rows] (w ~= hi + lo, 2^-17 relative), rows padded
  // to a multiple of 64 with zeros, so that one tensor map serves both parts and a k-block never straddles them.
  long long sWa, sWh, sK, sKp, sWdec, sWemb, stotal;                  // bf16 shadow offsets (elements); sKp: the LSTM kernel
                                                                      // again, columns interleaved for the fused gate kernel
  int pAtt, pK, pWdec, pWemb;                                         // shadow row pitches
  int rWa, rWh, rK, rWdec, rWemb;                                     // padded row counts (lo part starts at row r*)
};

static ParamLayout param_layout(bool gen, const sgg_dims_t& d) {
  ParamLayout L{};
  L.gen = gen; L.R = d.R; L.C = d.C; L.H = d.H; L.V = d.V; L.E = d.E;
  L.U = gen ? d.C : d.E;
  L.OUT = gen ? d.V : 1;
  L.KX = d.C + L.U + d.H;
  long long o = 0;
  auto take = [&](long long n) { long long r = o; o = rup(o + n, 64); return r; };
  L.Watt = take((long long)(d.R * d.C + d.H) * d.R);
  L.batt = take(d.R);
  L.K = take((long long)L.KX * 4 * d.H);
  for (int i = 0; i < 5; ++i) { L.lng[i] = take(d.H); L.lnb[i] = take(d.H); }
  L.Wdec = take((long long)d.H * L.OUT);
  L.bdec = take(L.OUT);
  L.Wemb = gen ? -1 : take((long long)d.V * d.E);
  L.total = o;
  long long s = 0;
  auto stake = [&](long long n) { long long r = s; s = rup(s + n, 128); return r; };
  // row pitches are multiples of 128 bytes: every 64-column TMA box row is then exactly one aligned L2 line
  // (W_a alone may instead be packed to an 8-element pitch -- 400 B rows for R = 196, 22 % fewer bytes per K1 pass but
  // rows that straddle L2 lines: SGG_WA_PITCH=dense selects it for A/B measurements)
  static int dense_att = -1;
  if (dense_att < 0) { const char* e = getenv("SGG_WA_PITCH"); dense_att = (e && e[0] == 'd') ? 1 : 0; }
  L.pAtt = dense_att ? (int)rup(d.R, 8) : (int)rup(d.R, 64); L.pK = 4 * d.H; L.pWdec = (int)rup(L.OUT, 64); L.pWemb = (int)rup(d.E, 64);
  L.rWa = (int)rup((long long)d.R * d.C, 64); L.rWh = (int)rup(d.H, 64); L.rK = (int)rup(L.KX, 64);
  L.rWdec = (int)rup(d.H, 64); L.rWemb = (int)rup(d.V, 64);
  L.sWa = stake(2LL * L.rWa * L.pAtt);
  L.sWh = stake(2LL * L.rWh * L.pAtt);
  L.sK = stake(2LL * L.rK * L.pK);
  L.sKp = stake(2LL * L.rK * L.pK);
  L.sWdec = gen ? stake(2LL * L.rWdec * L.pWdec) : -1;
  L.sWemb = gen ? -1 : stake(2LL * L.rWemb * L.pWemb);
  L.stotal = s;
  return L;
}

static const char* LN_NAMES[5] = {"input", "transform", "forget", "output", "state"};

static int fill_adam_segs(const ParamLayout& L, AdamSeg* seg) {
  int n = 0;
  const long long RC = (long long)L.R * L.C;
  seg[n++] = {L.Watt, L.sWa, L.R, L.pAtt, RC * L.R, (long long)L.rWa * L.pAtt};
  seg[n++] = {L.Watt + RC * L.R, L.sWh, L.R, L.pAtt, (long long)L.H * L.R, (long long)L.rWh * L.pAtt};
  seg[n++] = {L.batt, -1, L.R, L.R, L.R, 0};
  seg[n++] = {L.K, L.sK, 4 * L.H, L.pK, (long long)L.KX * 4 * L.H, (long long)L.rK * L.pK};
  for (int i = 0; i < 5; ++i) {
    seg[n++] = {L.lng[i], -1, L.H, L.H, L.H, 0};
    seg[n++] = {L.lnb[i], -1, L.H, L.H, L.H, 0};
  }
  seg[n++] = {L.Wdec, L.sWdec, L.OUT, L.pWdec, (long long)L.H * L.OUT, (long long)L.rWdec * L.pWdec};
  seg[n++] = {L.bdec, -1, L.OUT, L.OUT, L.OUT, 0};
  if (!L.gen) seg[n++] = {L.Wemb, L.sWemb, L.E, L.pWemb, (long long)L.V * L.E, (long long)L.rWemb * L.pWemb};
  for (int i = 0; i < n; ++i) seg[i].sh2_off = (seg[i].off == L.K) ? L.sKp : -1;
  return n;
}

// ============================================================================ workspace layout
struct NetWs {
  int NRmax;
  __nv_bfloat16* X; float* Cf; __nv_bfloat16* CH; float* EA; float* ED; float* Q; float* P;
  __nv_bfloat16* QB; float* XB; __nv_bfloat16* EB; float* CB; float* PB; __nv_bfloat16* PBH;
  float* Y;
  float* GSCR;   // exchange scratch of the fused gate kernel (gates.cu)
};
struct Ws {
  NetWs g, d;
  __nv_bfloat16* FAKE;   // [FS][T*B, 2*VP] generator logits hi/lo (row t*B+b), one slot per noise draw
  float* DFAKE; __nv_bfloat16* DFAKEH;  // [T*B, VP] fp32 + hi/lo : d gen_cost / d fake, or the GP gradient g
  __nv_bfloat16* VHL;    // [T*B, 2*VP] v = coef * g hi/lo
  float* HB;             // [T*B, H]
  float* UF;             // [T*B, EP] fp32 embedding temp
  float* UF2;            // ... of the tangent pass (so that both can be cleared by the step's one zero-fill launch)
  __nv_bfloat16* UBH;    // [T*B, 2*EP]
  __nv_bfloat16* UDB;    // [T*B, 2*EP]
  __nv_bfloat16* TRIH;   // [T*B, 2*VP] arbitrary float triples hi/lo (sgg_disc_forward)
  float* slopes; float* coef;
  float* LNP;            // [lstm_rev grid][LR_NPART] partial LN / head gradients
  float* LNP2;           // ... of the auxiliary stream of the two-stream reverse pass
  long long bytes;
};

constexpr int MAX_GEN_STREAMS = 8;   // noise draws served by one generator forward (= attention streams per tile read)
struct Dm {  // derived dimensions
  int B, T, V, R, C, H, E, RP, VP, EP, KXG, KXD, GS, FS;
};
static Dm derive(const sgg_dims_t& d) {
  Dm m{d.B, d.T, d.V, d.R, d.C, d.H, d.E, 0, 0, 0, 0, 0, 1, 1};
  m.FS = d.S > 1 ? d.S : 1;                                   // fake-logit slots kept in the workspace
  m.GS = m.FS < MAX_GEN_STREAMS ? m.FS : MAX_GEN_STREAMS;     // generator rows = GS * B
  m.RP = (int)rup(d.R, 64); m.VP = (int)rup(d.V, 64); m.EP = (int)rup(d.E, 64);
  m.KXG = (int)rup(d.C + d.C + d.H, 64); m.KXD = (int)rup(d.C + d.E + d.H, 64);
  return m;
}

static Ws ws_layout(const sgg_dims_t& d, void* base) {
  const Dm m = derive(d);
  Ws w{};
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  long long o = 0;
  auto take = [&](long long bytes) { uint8_t* r = p ? p + o : nullptr; o = rup(o + bytes, 256); return (void*)r; };
  auto net = [&](NetWs& n, int NR, int KXP, bool disc) {
    n.NRmax = NR;
    const long long T = m.T;
    n.X = (__nv_bfloat16*)take((T + 1) * NR * 2LL * KXP * 2);
    n.Cf = (float*)take((T + 1) * NR * (long long)m.H * 4);
    n.CH = (__nv_bfloat16*)take((T + 1) * NR * 2LL * m.H * 2);
    n.EA = (float*)take(T * NR * (long long)m.RP * 4);
    n.ED = disc ? (float*)take(T * (long long)m.B * m.RP * 4) : nullptr;
    n.Q = (float*)take(T * NR * 4LL * m.H * 4);
    n.P = (float*)take((long long)m.B * m.RP * 4);
    n.QB = (__nv_bfloat16*)take(T * NR * 8LL * m.H * 2);
    n.XB = (float*)take(T * NR * (long long)KXP * 4);
    n.EB = (__nv_bfloat16*)take(T * NR * 2LL * m.RP * 2);
    n.CB = (float*)take(T * NR * (long long)m.H * 4);
    n.PB = (float*)take((long long)m.B * m.RP * 4);
    n.PBH = (__nv_bfloat16*)take((long long)m.B * 2 * m.RP * 2);
    n.Y = disc ? (float*)take((long long)NR * T * 4) : nullptr;
    n.GSCR = (float*)take(gates_scratch_floats(NR) * 4);
  };
  net(w.g, m.GS * m.B, m.KXG, false);
  net(w.d, 4 * m.B, m.KXD, true);
  const long long TB = (long long)m.T * m.B;
  w.FAKE = (__nv_bfloat16*)take(m.FS * TB * 2 * m.VP * 2);
  w.DFAKE = (float*)take(TB * m.VP * 4);
  w.DFAKEH = (__nv_bfloat16*)take(TB * 2 * m.VP * 2);
  w.VHL = (__nv_bfloat16*)take(TB * 2 * m.VP * 2);
  w.HB = (float*)take(TB * m.H * 4);
  w.UF = (float*)take(TB * m.EP * 4);
  w.UF2 = (float*)take(TB * m.EP * 4);
  w.UBH = (__nv_bfloat16*)take(TB * 2 * m.EP * 2);
  w.UDB = (__nv_bfloat16*)take(TB * 2 * m.EP * 2);
  w.TRIH = (__nv_bfloat16*)take(TB * 2 * m.VP * 2);
  w.slopes = (float*)take(m.B * 4);
  w.coef = (float*)take(m.B * 4);
  w.LNP = (float*)take(lstm_rev_partials_floats() * 4);
  w.LNP2 = (float*)take(lstm_rev_partials_floats() * 4);
  w.bytes = o;
  return w;
}

// ============================================================================ one network
struct Net {
  bool gen;
  Dm m;
  ParamLayout L;
  const float* theta; const __nv_bfloat16* sh; float* grad;  // grad may be null (data path only)
  const __nv_bfloat16* a;
  NetWs w;
  float* lnp;    // per-CTA partial LN / head gradients of the reverse pass
  int NR;        // active rows per timestep (streams * B)
  int KXP, U, uoff, hoff;
  cudaStream_t st;
  // Forward-only sampling keeps no history: state / input buffers alternate between two slots and the per-step
  // scratch (scores, gate pre-activations) is reused, so the working set of a chunk stays in L2.
  bool roll = false;
  long long t2(int t) const { return roll ? (t & 1) : t; }
  long long t1(int t) const { return roll ? 0 : t; }
  // strides (elements) between timesteps, for the active NR
  long long sX() const { return (long long)NR * 2 * KXP; }
  long long sCf() const { return (long long)NR * m.H; }
  long long sCH() const { return (long long)NR * 2 * m.H; }
  long long sEA() const { return (long long)NR * m.RP; }
  long long sQ() const { return (long long)NR * 4 * m.H; }
  long long sQB() const { return (long long)NR * 8 * m.H; }
  long long sXB() const { return (long long)NR * KXP; }
  long long sEB() const { return (long long)NR * 2 * m.RP; }
  LstmLN ln() const {
    LstmLN l;
    for (int i = 0; i < 5; ++i) { l.gamma[i] = theta + L.lng[i]; l.beta[i] = theta + L.lnb[i]; }
    return l;
  }
};

static Net make_net(bool gen, const sgg_dims_t& d, const float* theta, const void* sh, float* grad, const void* a,
                    const NetWs& w, int NR, cudaStream_t st, float* lnp = nullptr) {
  Net n{};
  n.lnp = lnp;
  n.gen = gen; n.m = derive(d); n.L = param_layout(gen, d);
  n.theta = theta; n.sh = (const __nv_bfloat16*)sh; n.grad = grad; n.a = (const __nv_bfloat16*)a;
  n.w = w; n.NR = NR; n.st = st;
  n.KXP = gen ? n.m.KXG : n.m.KXD;
  n.U = n.L.U; n.uoff = d.C; n.hoff = d.C + n.L.U;
  return n;
}

// Three bf16 products of an fp32-faithful contraction  (x_hi + x_lo)(w_hi + w_lo) ~= x_hi w_hi + x_lo w_hi + x_hi w_lo.
// Activation x: [rows, 2*a_lo] with the lo part at column a0 + a_lo.  Weight (shadow): hi rows [0, r), lo rows [r, 2r).
//   weight used MN-major (forward, TF [in,out] layout): its rows are the contraction index  -> lo at k  offset b_lo
//   weight used K-major  (backward, x_bar = y_bar W^T): its rows are the output index       -> lo at mn offset b_lo
static void segs_act_weight(sgg_gemm_desc_t& g, int a0, int a_lo, int b_lo, bool weight_rows_are_k, int klen) {
  g.nseg = 3;
  for (int s = 0; s < 3; ++s) { g.seg_klen[s] = klen; g.seg_a_k[s] = a0; g.seg_b_k[s] = 0; g.seg_a_mn[s] = 0; g.seg_b_mn[s] = 0; }
  g.seg_a_k[1] = a0 + a_lo;
  if (weight_rows_are_k) g.seg_b_k[2] = b_lo; else g.seg_b_mn[2] = b_lo;
}

static sgg_gemm_desc_t gd_zero() { sgg_gemm_desc_t g; memset(&g, 0, sizeof(g)); g.alpha = 1.0f; return g; }
int gemm(const sgg_gemm_desc_t& d, cudaStream_t stream);
struct AdamProjParams;
int adam_proj(const AdamProjParams& p, const void* ann, cudaStream_t stream);

// K1: P = flat(a) W_a  (bias is added where P is consumed); hoisted out of the time loop (gen:14-15).
// lane: communicator lane (= stream) of the sharded variant's reduce-scatter.
static int net_attn_proj(const Net& n, int lane = 0) {
  const Dm& m = n.m;
  sgg_gemm_desc_t g = gd_zero();
  const long long K = (long long)m.R * m.C;
  if (t_shard) {   // partial projection of the GLOBAL batch over this rank's contraction slice, then reduce-scatter
    const ShardCtx& sc = *t_shard;
    const long long Bg = (long long)m.B * sc.world, k0 = sc.rank * sc.Ks;
    g.A = n.gen ? sc.slab_g : sc.slab_d; g.a_rows = Bg; g.a_cols = sc.Ks; g.a_ld = sc.Ks; g.a_mn_major = 0;
    g.B = n.sh + n.L.sWa; g.b_rows = 2LL * n.L.rWa; g.b_cols = m.R; g.b_ld = n.L.pAtt; g.b_mn_major = 1;
    g.M = (int)Bg; g.N = m.R; g.nseg = 2; g.seg_klen[0] = g.seg_klen[1] = (int)sc.Ks;
    g.seg_b_k[0] = (int)k0; g.seg_b_k[1] = n.L.rWa + (int)k0;
    float* part = sc.Ppart[n.gen ? 0 : 1];
    g.C = part; g.ldc = m.RP;
    SGG_TRY(gemm(g, n.st));
    return comm_reduce_scatter(sc.comm, part, n.w.P, (long long)m.B * m.RP, n.st, lane);
  }
  g.A = n.a; g.a_rows = m.B; g.a_cols = K; g.a_ld = K; g.a_mn_major = 0;
  g.B = n.sh + n.L.sWa; g.b_rows = 2LL * n.L.rWa; g.b_cols = m.R; g.b_ld = n.L.pAtt; g.b_mn_major = 1;
  g.M = m.B; g.N = m.R; g.nseg = 2; g.seg_klen[0] = g.seg_klen[1] = (int)K; g.seg_b_k[1] = n.L.rWa;  // a is bf16-exact
  g.C = n.w.P; g.ldc = m.RP;
  return gemm(g, n.st);
}

// c0 = h0 = mean_r a (gen:76-77) replicated into `nblk` stream blocks of step 0.
static int net_init_state(const Net& n, int nblk) {
  MeanPoolParams p{};
  p.a = n.a; p.B = n.m.B; p.R = n.m.R; p.nblk = nblk;
  p.C0 = n.w.Cf; p.CH = n.w.CH; p.ldCH = 2 * n.m.H; p.ch_lo = n.m.H;
  p.X = n.w.X; p.ldX = 2 * n.KXP; p.x_lo = n.KXP; p.hoff = n.hoff;
  return meanpool(p, n.st);
}

// e = P + b + c W_h for rows [row0, row0+nrows) of step t (tangent: edot = cdot W_h into ED).
static int net_scores(const Net& n, int t, int row0, int nrows, bool tangent, bool prezeroed = false) {
  const Dm& m = n.m;
  sgg_gemm_desc_t g = gd_zero();
  g.A = n.w.CH + n.t2(t) * n.sCH() + (long long)row0 * 2 * m.H; g.a_rows = nrows; g.a_cols = 2 * m.H; g.a_ld = 2 * m.H;
  g.B = n.sh + n.L.sWh; g.b_rows = 2LL * n.L.rWh; g.b_cols = m.R; g.b_ld = n.L.pAtt; g.b_mn_major = 1;
  g.M = nrows; g.N = m.R;
  segs_act_weight(g, 0, m.H, n.L.rWh, true, m.H);
  if (tangent) {
    g.C = n.w.ED + (long long)t * m.B * m.RP; g.ldc = m.RP;
  } else {
    g.C = n.w.EA + n.t1(t) * n.sEA() + (long long)row0 * m.RP; g.ldc = m.RP;
    g.bias = n.theta + n.L.batt;
    g.addm = n.w.P; g.ld_addm = m.RP; g.add_mod = m.B;   // row0 is a multiple of B
  }
  if (prezeroed) g.atomic = 2;   // the preceding cell kernel cleared these rows (ZeroRow)
  return gemm(g, n.st);
}

// q = [z,u,h] K for rows [row0, row0+nrows) of step t.
static int net_gates(const Net& n, int t, int row0, int nrows, bool prezeroed = false) {
  const Dm& m = n.m;
  sgg_gemm_desc_t g = gd_zero();
  g.A = n.w.X + n.t2(t) * n.sX() + (long long)row0 * 2 * n.KXP; g.a_rows = nrows; g.a_cols = 2 * n.KXP; g.a_ld = 2 * n.KXP;
  g.B = n.sh + n.L.sK; g.b_rows = 2LL * n.L.rK; g.b_cols = 4 * m.H; g.b_ld = n.L.pK; g.b_mn_major = 1;
  g.M = nrows; g.N = 4 * m.H;
  segs_act_weight(g, 0, n.KXP, n.L.rK, true, n.KXP);
  g.C = n.w.Q + n.t1(t) * n.sQ() + (long long)row0 * 4 * m.H; g.ldc = 4 * m.H;
  if (prezeroed) g.atomic = 2;   // the attention kernel in front cleared these rows
  return gemm(g, n.st);
}

// One primal forward step (scores, attention, gates, cell) for stream blocks [0, nblk).
static int net_forward_step(const Net& n, int t, int nblk, bool ea0_prezeroed = false) {
  const Dm& m = n.m;
  const int rows = nblk * m.B;
  const bool fused = gates_fused_available() && n.w.GSCR != nullptr;
  SGG_TRY(net_scores(n, t, 0, rows, false, /*prezeroed=*/t > 0 || ea0_prezeroed));
  AttnFwdParams ap{};
  ap.a = n.a; ap.B = m.B; ap.R = m.R; ap.nv = nblk; ap.early_a = t_ann_static;
  for (int v = 0; v < nblk; ++v) { ap.row_blk[v] = v; ap.e_blk[v] = v; }
  ap.E = n.w.EA + n.t1(t) * n.sEA(); ap.ldE = m.RP;
  ap.alpha_out = n.w.EA + n.t1(t) * n.sEA(); ap.ldA = m.RP;
  ap.X = n.w.X + n.t2(t) * n.sX(); ap.ldX = 2 * n.KXP; ap.lo_off = n.KXP;
  if (!fused) { ap.zero_p = n.w.Q + n.t1(t) * n.sQ(); ap.zero_ld = 4 * m.H; ap.zero_cols = 4 * m.H; }
  SGG_TRY(attn_fwd(ap, 0, n.st));
  if (fused) {   // gate GEMM with the LayerNorm / cell epilogue fused in (gates.cu): one launch, no read-back of q
    GatesParams gp{};
    gp.nrows = rows;
    gp.Cin = n.w.Cf + n.t2(t) * n.sCf();
    gp.ln = n.ln();
    gp.Q = n.w.Q + n.t1(t) * n.sQ(); gp.ldQ = 4 * m.H;
    gp.Cout = n.w.Cf + n.t2(t + 1) * n.sCf();
    gp.CH = n.w.CH + n.t2(t + 1) * n.sCH(); gp.ldCH = 2 * m.H; gp.ch_lo = m.H;
    gp.Xn = n.w.X + n.t2(t + 1) * n.sX(); gp.ldX = 2 * n.KXP; gp.x_lo = n.KXP; gp.hoff = n.hoff;
    if (!n.gen) {
      gp.wdec = n.theta + n.L.Wdec; gp.bdec = n.theta + n.L.bdec;
      gp.Y = n.w.Y + t; gp.ldY = m.T;
    }
    if (t + 1 < m.T) gp.zero = ZeroRow{n.w.EA + n.t1(t + 1) * n.sEA(), m.RP, m.RP};
    gp.scratch = n.w.GSCR;
    return gates_fused(n.w.X + n.t2(t) * n.sX(), 2 * n.KXP, n.KXP, n.sh + n.L.sKp, n.L.rK, gp, n.st);
  }
  SGG_TRY(net_gates(n, t, 0, rows, true));
  LstmFwdParams lp{};
  if (t + 1 < m.T) lp.zero = ZeroRow{n.w.EA + n.t1(t + 1) * n.sEA(), m.RP, m.RP};
  lp.nrows = rows;
  lp.Q = n.w.Q + n.t1(t) * n.sQ(); lp.ldQ = 4 * m.H;
  lp.Cin = n.w.Cf + n.t2(t) * n.sCf();
  lp.ln = n.ln();
  lp.Cout = n.w.Cf + n.t2(t + 1) * n.sCf();
  lp.CH = n.w.CH + n.t2(t + 1) * n.sCH(); lp.ldCH = 2 * m.H; lp.ch_lo = m.H;
  lp.Xn = n.w.X + n.t2(t + 1) * n.sX(); lp.ldX = 2 * n.KXP; lp.x_lo = n.KXP; lp.hoff = n.hoff;
  if (!n.gen) {
    lp.wdec = n.theta + n.L.Wdec; lp.bdec = n.theta + n.L.bdec;
    lp.Y = n.w.Y + t; lp.ldY = m.T;
  }
  return lstm_fwd(lp, n.st);
}

// Primal forward over T steps for stream blocks [0, nblk).  Needs: X[t] u-columns filled, P, state 0.
static int net_forward(const Net& n, int nblk, bool ea0_prezeroed = false) {
  for (int t = 0; t < n.m.T; ++t) SGG_TRY(net_forward_step(n, t, nblk, ea0_prezeroed));
  return 0;
}

// Tangent forward (D only): tangent rows are block `tblk`, their primal partner block `pblk`.
static int net_tangent(const Net& n, int pblk, int tblk, bool q0_prezeroed = false) {
  const Dm& m = n.m;
  for (int t = 0; t < m.T; ++t) {
    if (t > 0) {  // cdot_0 = 0 => edot_0 = adot_0 = zdot_0 = 0 (z columns of X[0] tangent rows stay zero)
      SGG_TRY(net_scores(n, t, tblk * m.B, m.B, true, true));   // cleared by lstm_tan of step t-1
      AttnFwdParams ap{};
      ap.a = n.a; ap.B = m.B; ap.R = m.R; ap.nv = 1; ap.early_a = t_ann_static;
      ap.row_blk[0] = tblk; ap.e_blk[0] = 0; ap.ain_blk = pblk;
      ap.E = n.w.ED + (long long)t * m.B * m.RP; ap.ldE = m.RP;
      ap.alpha_in = n.w.EA + t * n.sEA();
      ap.alpha_out = n.w.EA + t * n.sEA(); ap.ldA = m.RP;
      ap.X = n.w.X + t * n.sX(); ap.ldX = 2 * n.KXP; ap.lo_off = n.KXP;
      ap.zero_p = n.w.Q + t * n.sQ(); ap.zero_ld = 4 * m.H; ap.zero_cols = 4 * m.H;
      SGG_TRY(attn_fwd(ap, 1, n.st));
    }
    SGG_TRY(net_gates(n, t, tblk * m.B, m.B, t > 0 || q0_prezeroed));
    LstmTanParams lp{};
    if (t + 1 < m.T) lp.zero = ZeroRow{n.w.ED + (long long)(t + 1) * m.B * m.RP, m.RP, m.RP};
    lp.nrows = m.B; lp.prow0 = pblk * m.B; lp.trow0 = tblk * m.B;
    lp.Q = n.w.Q + t * n.sQ(); lp.ldQ = 4 * m.H;
    lp.C = n.w.Cf + t * n.sCf();
    lp.ln = n.ln();
    lp.Cout = n.w.Cf + (t + 1) * n.sCf();
    lp.CH = n.w.CH + (t + 1) * n.sCH(); lp.ldCH = 2 * m.H; lp.ch_lo = m.H;
    lp.Xn = n.w.X + (t + 1) * n.sX(); lp.ldX = 2 * n.KXP; lp.x_lo = n.KXP; lp.hoff = n.hoff;
    SGG_TRY(lstm_tan(lp, n.st));
  }
  return 0;
}

struct RevCfg {
  int blk0, nblk;        // primal stream blocks [blk0, blk0+nblk)
  int tan_pblk, tan_blk; // tangent pairing (or -1)
  float ybar_blk[8];     // D head upstream per block
  float ydot_bar;        // D head tangent upstream (lambda)
  const float* HB;       // G: [T*B, H] h_bar contributions from the logits (row t*B+b), or null
  bool wgrad;            // accumulate parameter gradients
  bool pb_prezeroed;     // P_bar was cleared by the step's zero-fill list
  bool clear_P;          // the W_a chain's hi/lo packing kernel also clears P (the fused Adam + projection accumulates into it)
  bool ann_grad;         // the caller wants d cost / d annotations: c0 = mean_r a is then a function of an input, so the
                         // step-0 state adjoint is completed too (c_bar_0 += e_bar_0 W_h^T)
};

// Reverse time loop over T steps for the stream blocks of `rc` (with wgrad: LN / head gradient partials into n.lnp,
// P_bar accumulated by attn_rev into PB, which the caller has cleared).
static int net_reverse_loop(const Net& n, const RevCfg& rc) {
  const Dm& m = n.m;
  const bool tan = rc.tan_blk >= 0;
  const int row0 = rc.blk0 * m.B;
  const int nrows_p = rc.nblk * m.B;                       // primal rows
  const int nrows_all = nrows_p + (tan ? m.B : 0);         // tangent block directly follows
  for (int t = m.T - 1; t >= 0; --t) {
    const bool last = (t == m.T - 1);
    LstmRevParams lp{};
    lp.B = m.B;
    lp.Q = n.w.Q + t * n.sQ(); lp.ldQ = 4 * m.H;
    lp.C = n.w.Cf + t * n.sCf();
    lp.ln = n.ln();
    lp.XBn = last ? nullptr : n.w.XB + (t + 1) * n.sXB(); lp.ldXB = n.KXP; lp.hoff = n.hoff;
    lp.HB = rc.HB ? rc.HB + (long long)t * m.B * m.H : nullptr; lp.ldHB = m.H;
    lp.CBn = last ? nullptr : n.w.CB + (t + 1) * n.sCf();
    for (int i = 0; i < 8; ++i) lp.ybar_blk[i] = rc.ybar_blk[i];
    lp.ydot_bar = rc.ydot_bar;
    lp.wdec = n.gen ? nullptr : n.theta + n.L.Wdec;
    lp.QB = n.w.QB + t * n.sQB(); lp.ldQB = 8 * m.H; lp.qb_lo = 4 * m.H;
    lp.CB = n.w.CB + t * n.sCf();
    if (rc.wgrad) { lp.partials = n.lnp; lp.init_partials = last ? 1 : 0; }
    lp.zero = ZeroRow{n.w.XB + t * n.sXB(), n.KXP, n.KXP};   // x_bar rows of this step: the GEMM below accumulates into them
    // first-order rows, then the (interp, tangent) pair, in one launch
    lp.n_plain = (tan ? (rc.tan_pblk - rc.blk0) : rc.nblk) * m.B; lp.prow0 = row0;
    lp.n_tan = tan ? m.B : 0; lp.tan_prow0 = tan ? rc.tan_pblk * m.B : 0; lp.trow0 = tan ? rc.tan_blk * m.B : 0;
    SGG_TRY(lstm_rev(lp, n.st));
    // x_bar = q_bar K^T  (rows incl. tangent)
    {
      sgg_gemm_desc_t g = gd_zero();
      g.A = n.w.QB + t * n.sQB() + (long long)row0 * 8 * m.H; g.a_rows = nrows_all; g.a_cols = 8 * m.H; g.a_ld = 8 * m.H;
      g.B = n.sh + n.L.sK; g.b_rows = 2LL * n.L.rK; g.b_cols = 4 * m.H; g.b_ld = n.L.pK; g.b_mn_major = 0;
      g.M = nrows_all; g.N = n.L.KX;
      segs_act_weight(g, 0, 4 * m.H, n.L.rK, false, 4 * m.H);
      g.C = n.w.XB + t * n.sXB() + (long long)row0 * n.KXP; g.ldc = n.KXP; g.atomic = 2;
      SGG_TRY(gemm(g, n.st));
    }
    if (t == 0 && !rc.wgrad) break;  // data path: nothing upstream of the step-0 attention is needed
    AttnRevParams ap{};
    ap.a = n.a; ap.B = m.B; ap.R = m.R; ap.early_a = t_ann_static;
    ap.nv = rc.nblk + (tan ? 1 : 0);
    ap.tan_stream = tan ? (rc.tan_pblk - rc.blk0) : -1;
    for (int v = 0; v < rc.nblk; ++v) ap.row_blk[v] = rc.blk0 + v;
    if (tan) ap.row_blk[rc.nblk] = rc.tan_blk;
    ap.XB = n.w.XB + t * n.sXB(); ap.ldXB = n.KXP;
    ap.alpha = n.w.EA + t * n.sEA(); ap.ldA = m.RP;
    ap.edot = tan ? n.w.ED + (long long)t * m.B * m.RP : nullptr;
    ap.EB = n.w.EB + t * n.sEB(); ap.ldEB = 2 * m.RP; ap.lo_off = m.RP;
    ap.Pbar = rc.wgrad ? n.w.PB : nullptr; ap.ldP = m.RP;
    SGG_TRY(attn_rev(ap, n.st));
    if (t > 0 || rc.ann_grad) {  // c_bar of step t (in place) += e_bar W_h^T
      sgg_gemm_desc_t g = gd_zero();
      g.A = n.w.EB + t * n.sEB() + (long long)row0 * 2 * m.RP; g.a_rows = nrows_all; g.a_cols = 2 * m.RP; g.a_ld = 2 * m.RP;
      g.B = n.sh + n.L.sWh; g.b_rows = 2LL * n.L.rWh; g.b_cols = m.R; g.b_ld = n.L.pAtt; g.b_mn_major = 0;
      g.M = nrows_all; g.N = m.H;
      segs_act_weight(g, 0, m.RP, n.L.rWh, false, m.RP);
      float* cb = n.w.CB + t * n.sCf() + (long long)row0 * m.H;
      g.C = cb; g.ldc = m.H; g.atomic = 1;   // c_bar += e_bar W_h^T
      SGG_TRY(gemm(g, n.st));
    }
  }
  return 0;
}
// CTAs (= partial slices) the lstm_rev launches of a reverse loop over `rc` use
static int rev_slices(const Net& n, const RevCfg& rc) {
  const bool tan = rc.tan_blk >= 0;
  return lstm_rev_grid((tan ? (rc.tan_pblk - rc.blk0) : rc.nblk) * n.m.B + (tan ? n.m.B : 0));
}

// Tiny reduction kernels (losses, LN-gradient partial sums, column sums, the one-hot embedding scatter) occupy a handful of
// CTAs for a few microseconds each; on the step's serial chain each of them costs a full launch slot.  They run on the
// auxiliary stream instead (forked where their inputs are complete, joined at the end of the step), beside the SM-filling
// kernels of the main chain.  SGG_AUX_STREAM=0 keeps them in line.
static bool aux_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SGG_AUX_STREAM"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1 && side_stream(1) != nullptr;
}
static int aux_fork(cudaStream_t st, cudaStream_t* out) {
  *out = st;
  if (!aux_enabled()) return 0;
  return side_fork(st, out, 1);
}
static int aux_join(cudaStream_t st) {
  if (!aux_enabled()) return 0;
  return side_join(st, side_stream(1)->s, 1);
}

// Parameter gradients after the reverse loop(s) over ALL active rows of the network: LN gamma / beta (and D head) from
// the per-CTA partials of up to two loops, then one GEMM per kernel over all timesteps / streams.
static int net_reverse_wgrad(const Net& n, bool clear_P, cudaStream_t* side_out, const float* lnp0, int slices0,
                             const float* lnp1 = nullptr, int slices1 = 0) {
  const Dm& m = n.m;
  cudaStream_t ax = n.st;
  SGG_TRY(aux_fork(n.st, &ax));   // joined by the caller at the end of the step (aux_join)
  for (int k = 0; k < 2; ++k) {
    const float* part = k == 0 ? lnp0 : lnp1;
    if (!part) continue;
    LnGradParams lg{};
    lg.partials = part; lg.nslices = k == 0 ? slices0 : slices1;
    for (int i = 0; i < 5; ++i) { lg.dgamma[i] = n.grad + n.L.lng[i]; lg.dbeta[i] = n.grad + n.L.lnb[i]; }
    if (!n.gen) { lg.dwdec = n.grad + n.L.Wdec; lg.dbdec = n.grad + n.L.bdec; }
    SGG_TRY(lngrad_reduce(lg, ax));
  }
  // ---------------- weight gradients: one GEMM per kernel over all timesteps / streams
  const long long rowsT = (long long)m.T * n.NR;
  SGG_TRY(colsum(n.w.PB, m.RP, m.B, m.R, n.grad + n.L.batt, ax));   // db_att = column sums of P_bar
  {  // dW_a = flat(a)^T P_bar  [R*C, R]: on the side stream when the caller takes it over (it joins later)
    cudaStream_t s1 = n.st;
    if (side_out) { SGG_TRY(side_fork(n.st, &s1)); *side_out = s1; }
    const int lane = s1 != n.st ? 1 : 0;
    const long long K = (long long)m.R * m.C;
    PackParams pk{};
    sgg_gemm_desc_t g = gd_zero();
    if (t_shard) {   // rows [k0, k0+Ks) of dW_a over the GLOBAL batch: all-gather P_bar, contract with this rank's slab
      const ShardCtx& sc = *t_shard;
      const long long Bg = (long long)m.B * sc.world, k0 = sc.rank * sc.Ks;
      float* pball = sc.PBall[n.gen ? 0 : 1];
      __nv_bfloat16* pbhall = sc.PBHall[n.gen ? 0 : 1];
      SGG_TRY(comm_all_gather(sc.comm, n.w.PB, pball, (long long)m.B * m.RP, s1, lane));
      pk.rows = (int)Bg; pk.cols = m.R; pk.src = pball; pk.ld = m.RP;
      pk.dst = pbhall; pk.ldd = 2 * m.RP; pk.lo_off = m.RP;
      SGG_TRY(pack_hl(pk, s1));
      g.A = n.gen ? sc.slab_g : sc.slab_d; g.a_rows = Bg; g.a_cols = sc.Ks; g.a_ld = sc.Ks; g.a_mn_major = 1;
      g.B = pbhall; g.b_rows = Bg; g.b_cols = 2 * m.RP; g.b_ld = 2 * m.RP; g.b_mn_major = 1;
      g.M = (int)sc.Ks; g.N = m.R; g.nseg = 2;
      g.seg_klen[0] = g.seg_klen[1] = (int)Bg; g.seg_b_mn[1] = m.RP;
      g.C = n.grad + n.L.Watt + k0 * m.R; g.ldc = m.R; g.atomic = 0; g.splits = 0;
      SGG_TRY(gemm(g, s1));
    } else {
      pk.rows = m.B; pk.cols = m.R; pk.src = n.w.PB; pk.ld = m.RP;
      pk.dst = n.w.PBH; pk.ldd = 2 * m.RP; pk.lo_off = m.RP;
      if (clear_P) { pk.zero_p = reinterpret_cast<float4*>(n.w.P); pk.zero_n4 = (long long)m.B * m.RP / 4; }
      SGG_TRY(pack_hl(pk, s1));
      g.A = n.a; g.a_rows = m.B; g.a_cols = K; g.a_ld = K; g.a_mn_major = 1;
      g.B = n.w.PBH; g.b_rows = m.B; g.b_cols = 2 * m.RP; g.b_ld = 2 * m.RP; g.b_mn_major = 1;
      g.M = (int)K; g.N = m.R; g.nseg = 2;
      g.seg_klen[0] = g.seg_klen[1] = m.B; g.seg_b_mn[1] = m.RP;
      g.C = n.grad + n.L.Watt; g.ldc = m.R; g.atomic = 0; g.splits = 1;
      SGG_TRY(gemm(g, s1));
    }
  }
  {  // dK = X^T QB   [KX, 4H], three hi/lo products
    sgg_gemm_desc_t g = gd_zero();
    g.A = n.w.X; g.a_rows = rowsT; g.a_cols = 2 * n.KXP; g.a_ld = 2 * n.KXP; g.a_mn_major = 1;
    g.B = n.w.QB; g.b_rows = rowsT; g.b_cols = 8 * m.H; g.b_ld = 8 * m.H; g.b_mn_major = 1;
    g.M = n.L.KX; g.N = 4 * m.H; g.nseg = 3;
    for (int s = 0; s < 3; ++s) g.seg_klen[s] = (int)rowsT;
    g.seg_b_mn[1] = 4 * m.H; g.seg_a_mn[2] = n.KXP;
    g.C = n.grad + n.L.K; g.ldc = 4 * m.H; g.atomic = 1;
    SGG_TRY(gemm(g, n.st));
  }
  {  // dW_h = C^T EB   [H, R]
    sgg_gemm_desc_t g = gd_zero();
    g.A = n.w.CH; g.a_rows = rowsT; g.a_cols = 2 * m.H; g.a_ld = 2 * m.H; g.a_mn_major = 1;
    g.B = n.w.EB; g.b_rows = rowsT; g.b_cols = 2 * m.RP; g.b_ld = 2 * m.RP; g.b_mn_major = 1;
    g.M = m.H; g.N = m.R; g.nseg = 3;
    for (int s = 0; s < 3; ++s) g.seg_klen[s] = (int)rowsT;
    g.seg_b_mn[1] = m.RP; g.seg_a_mn[2] = m.H;
    g.C = n.grad + n.L.Watt + (long long)m.R * m.C * m.R; g.ldc = m.R; g.atomic = 1;
    SGG_TRY(gemm(g, n.st));
  }
  return 0;
}

// Reverse pass over T steps as one sequence: loop, then (with wgrad) the parameter gradients.
static int net_reverse(const Net& n, const RevCfg& rc, cudaStream_t* side_out = nullptr) {
  const Dm& m = n.m;
  if (rc.wgrad && !rc.pb_prezeroed) SGG_TRY(zero_2d(n.w.PB, (long long)m.B * m.RP, (long long)m.B * m.RP, 1, n.st));
  SGG_TRY(net_reverse_loop(n, rc));
  if (!rc.wgrad) return 0;
  const bool tan = rc.tan_blk >= 0;
  SGG_CHECK(rc.blk0 == 0 && rc.nblk * m.B + (tan ? m.B : 0) == n.NR, "net_reverse: weight gradients need all active rows");
  return net_reverse_wgrad(n, rc.clear_P, side_out, n.lnp, rev_slices(n, rc));
}

// d cost / d annotations [B,R,C] fp32 (gen:68 / disc:68 self.downsampled) after a reverse pass with weight gradients and
// RevCfg::ann_grad over primal blocks [0, nprimal) (+ the tangent block tan_blk): K1's reverse P_bar W_a^T as a GEMM that
// overwrites `out`, then the attention / initial-state terms (misc.cu ann_grad_kernel).  Needs P_bar packed hi/lo (done
// by the dW_a chain) and the replicated W_a shadow (not available under the row-sharded projection).
struct AnnGradParams;
int ann_grad(const AnnGradParams& p, cudaStream_t stream);
static int net_ann_grad(const Net& n, int nprimal, int tan_blk, float* out) {
  const Dm& m = n.m;
  SGG_CHECK(t_shard == nullptr, "annotation gradients are not available with the row-sharded attention projection");
  SGG_CHECK(nprimal + (tan_blk >= 0 ? 1 : 0) <= 8, "net_ann_grad: too many stream blocks");
  const long long K = (long long)m.R * m.C;
  sgg_gemm_desc_t g = gd_zero();
  g.A = n.w.PBH; g.a_rows = m.B; g.a_cols = 2 * m.RP; g.a_ld = 2 * m.RP;
  g.B = n.sh + n.L.sWa; g.b_rows = 2LL * n.L.rWa; g.b_cols = m.R; g.b_ld = n.L.pAtt; g.b_mn_major = 0;
  g.M = m.B; g.N = (int)K;
  segs_act_weight(g, 0, m.RP, n.L.rWa, false, m.RP);
  g.C = out; g.ldc = K; g.splits = 1;
  SGG_TRY(gemm(g, n.st));
  AnnGradParams p{};
  p.B = m.B; p.R = m.R; p.T = m.T; p.nv = nprimal + (tan_blk >= 0 ? 1 : 0);
  for (int v = 0; v < nprimal; ++v) p.row_blk[v] = v;
  p.tan_v = -1;
  if (tan_blk >= 0) { p.row_blk[nprimal] = tan_blk; p.tan_v = nprimal; }
  p.alpha = n.w.EA; p.ldA = m.RP; p.strideA = n.sEA();
  p.XB = n.w.XB; p.ldXB = n.KXP; p.strideXB = n.sXB(); p.hoff = n.hoff;
  p.CB0 = n.w.CB;
  p.out = out;
  return ann_grad(p, n.st);
}

// ============================================================================ generator forward
static __nv_bfloat16* fake_slot(const Ws& w, const Dm& m, int slot) {
  return w.FAKE + (long long)slot * m.T * m.B * 2 * m.VP;
}

// Generator forward (gen:74-91) for `ns` noise draws at once: rows s*B+b of every buffer belong to draw s, and
// the ns streams of sample b share one read of its annotation tile.  g.NR must be ns*B.  Logits of draw s go
// to FAKE slot slot0+s (hi/lo, rows t*B+b).  noise: [ns, B, C] fp32.
static int gen_forward(const Net& g, const Ws& w, const float* noise, int ns, int slot0, bool recompute_proj,
                       float* logits_out) {
  const Dm& m = g.m;
  SGG_CHECK(ns >= 1 && ns <= m.GS && g.NR == ns * m.B && slot0 + ns <= m.FS, "gen_forward: bad stream count %d", ns);
  if (recompute_proj) {
    SGG_TRY(net_attn_proj(g));
    SGG_TRY(net_init_state(g, m.GS));   // every stream block of step 0 (any later ns <= GS finds its state)
  }
  {  // u_t = noise for every t (gen:81,86): hi/lo into the u columns of X[0..T-1]
    PackParams pk{};
    pk.rows = ns * m.B; pk.cols = m.C; pk.src = noise; pk.ld = m.C;
    pk.dst = g.w.X + g.uoff; pk.ldd = 2 * g.KXP; pk.lo_off = g.KXP;
    pk.reps = m.T; pk.rep_stride = g.sX();
    SGG_TRY(pack_hl(pk, g.st));
  }
  SGG_TRY(net_forward(g, ns));
  // logits for all timesteps and streams in one GEMM: A rows t*NR + s*B + b (h_{t+1} lives in X[t+1], gen:88),
  // stored stream-major: slot (slot0+s), row t*B+b
  sgg_gemm_desc_t d = gd_zero();
  d.A = g.w.X + g.sX(); d.a_rows = (long long)m.T * g.NR; d.a_cols = 2 * g.KXP; d.a_ld = 2 * g.KXP;
  d.B = g.sh + g.L.sWdec; d.b_rows = 2LL * g.L.rWdec; d.b_cols = m.V; d.b_ld = g.L.pWdec; d.b_mn_major = 1;
  d.M = m.T * g.NR; d.N = m.V;
  segs_act_weight(d, g.hoff, g.KXP, g.L.rWdec, true, m.H);
  d.bias = g.theta + g.L.bdec;
  d.Chl = fake_slot(w, m, slot0); d.ld_hl = 2 * m.VP; d.lo_off = m.VP;
  d.out_d0 = g.NR; d.out_d1 = m.B; d.out_s0 = m.B; d.out_s1 = (long long)m.T * m.B;
  SGG_TRY(gemm(d, g.st));
  if (logits_out) {  // [B,T,V] fp32 for the caller (stream 0 only): one strided store per timestep
    for (int t = 0; t < m.T; ++t) {
      sgg_gemm_desc_t e = d;
      e.A = g.w.X + (t + 1) * g.sX(); e.a_rows = m.B; e.M = m.B;
      e.out_d0 = 0;
      e.Chl = nullptr; e.C = logits_out + (long long)t * m.V; e.ldc = (long long)m.T * m.V;
      SGG_TRY(gemm(e, g.st));
    }
  }
  return 0;
}

// u = x W_emb for a dense [T*B, 2*VP] hi/lo input; result fp32 in w.UF
static int embed_dense(const Net& d, const Ws& w, const __nv_bfloat16* xhl, bool prezeroed = false, float* out = nullptr) {
  const Dm& m = d.m;
  sgg_gemm_desc_t g = gd_zero();
  g.A = xhl; g.a_rows = (long long)m.T * m.B; g.a_cols = 2 * m.VP; g.a_ld = 2 * m.VP;
  g.B = d.sh + d.L.sWemb; g.b_rows = 2LL * d.L.rWemb; g.b_cols = m.E; g.b_ld = d.L.pWemb; g.b_mn_major = 1;
  g.M = m.T * m.B; g.N = m.E;
  segs_act_weight(g, 0, m.VP, d.L.rWemb, true, m.VP);
  g.C = out ? out : w.UF; g.ldc = m.EP;
  if (prezeroed) g.atomic = 2;   // the output was cleared by the caller's zero-fill list
  return gemm(g, d.st);
}

// g = u_bar W_emb^T for rows t*B+b, u_bar taken from XB[t][blk] ; result fp32 in w.DFAKE (+hi/lo)
static int embed_input_grad(const Net& d, const Ws& w, int blk, bool want_hl) {
  const Dm& m = d.m;
  PackParams pk{};
  pk.rows = m.T * m.B; pk.cols = m.E; pk.rpg = m.B;
  pk.src = d.w.XB + (long long)blk * m.B * d.KXP + d.uoff; pk.ld = d.KXP; pk.gstride = d.sXB();
  pk.dst = w.UBH; pk.ldd = 2 * m.EP; pk.lo_off = m.EP;
  SGG_TRY(pack_hl(pk, d.st));
  sgg_gemm_desc_t g = gd_zero();
  g.A = w.UBH; g.a_rows = (long long)m.T * m.B; g.a_cols = 2 * m.EP; g.a_ld = 2 * m.EP;
  g.B = d.sh + d.L.sWemb; g.b_rows = 2LL * d.L.rWemb; g.b_cols = m.E; g.b_ld = d.L.pWemb; g.b_mn_major = 0;
  g.M = m.T * m.B; g.N = m.V;
  segs_act_weight(g, 0, m.EP, d.L.rWemb, false, m.EP);
  g.C = w.DFAKE; g.ldc = m.VP;
  if (want_hl) { g.Chl = w.DFAKEH; g.ld_hl = 2 * m.VP; g.lo_off = m.VP; }
  return gemm(g, d.st);
}

}  // namespace sgg

// ============================================================================ C ABI
using namespace sgg;

static int check_dims(const sgg_dims_t& d) {
  SGG_CHECK(d.C == 512 && d.H == 512, "dims: C and H must be 512 (reference gen:68,79); got C=%d H=%d", d.C, d.H);
  SGG_CHECK(d.R >= 1 && d.R <= 256, "dims: R=%d out of range 1..256", d.R);
  SGG_CHECK(d.B >= 1 && d.T >= 1 && d.V >= 1 && d.E >= 1, "dims: non-positive B/T/V/E");
  return 0;
}

extern "C" int sgg_param_table(int net, const sgg_dims_t* d, sgg_param_entry_t* out, int max_entries,
                               int* n_entries, int64_t* n_floats, int64_t* n_shadow) {
  SGG_CHECK(d != nullptr, "sgg_param_table: null dims");
  SGG_TRY(check_dims(*d));
  const bool gen = net == 0;
  const ParamLayout L = param_layout(gen, *d);
  const char* pre = gen ? "Generator/Generator" : "Discriminator/Discriminator";
  int n = 0;
  auto add = [&](const char* suffix, long long off, int rows, int cols, long long soff, int pitch, bool full_name,
                 int srows = 0) {
    if (out && n < max_entries) {
      sgg_param_entry_t& e = out[n];
      memset(&e, 0, sizeof(e));
      if (full_name) snprintf(e.name, sizeof(e.name), "%s", suffix);
      else snprintf(e.name, sizeof(e.name), "%s/%s", pre, suffix);
      e.offset = off; e.rows = rows; e.cols = cols; e.shadow_offset = soff; e.shadow_pitch = pitch; e.shadow_rows = srows;
    }
    ++n;
  };
  char buf[96];
  // (the attention kernel's shadow is split into its annotation rows W_a at sWa and its state rows W_h at sWh)
  add("attention_perceptron/kernel", L.Watt, L.R * L.C + L.H, L.R, L.sWa, L.pAtt, false, L.rWa);
  add("attention_perceptron/bias", L.batt, 1, L.R, -1, 0, false);
  add("layer_norm_basic_lstm_cell/kernel", L.K, L.KX, 4 * L.H, L.sK, L.pK, false, L.rK);
  for (int i = 0; i < 5; ++i) {
    snprintf(buf, sizeof(buf), "layer_norm_basic_lstm_cell/%s/gamma", LN_NAMES[i]);
    add(buf, L.lng[i], 1, L.H, -1, 0, false);
    snprintf(buf, sizeof(buf), "layer_norm_basic_lstm_cell/%s/beta", LN_NAMES[i]);
    add(buf, L.lnb[i], 1, L.H, -1, 0, false);
  }
  add("decoder/kernel", L.Wdec, L.H, L.OUT, L.sWdec, L.pWdec, false, L.rWdec);
  add("decoder/bias", L.bdec, 1, L.OUT, -1, 0, false);
  if (!gen) add("Discriminator/W", L.Wemb, L.V, L.E, L.sWemb, L.pWemb, true, L.rWemb);
  if (n_entries) *n_entries = n;
  if (n_floats) *n_floats = L.total;
  if (n_shadow) *n_shadow = L.stotal;
  return 0;
}

extern "C" int64_t sgg_workspace_bytes(const sgg_dims_t* d) {
  if (!d || check_dims(*d) != 0) return -1;
  return ws_layout(*d, nullptr).bytes;
}

extern "C" int sgg_refresh_shadow(int net, const sgg_dims_t* d, const float* theta, void* shadow, sgg_stream_t stream) {
  SGG_CHECK(d && theta && shadow, "sgg_refresh_shadow: null argument");
  SGG_TRY(check_dims(*d));
  const ParamLayout L = param_layout(net == 0, *d);
  AdamSeg seg[ADAM_MAX_SEG];
  const int n = fill_adam_segs(L, seg);
  for (int i = 0; i < n; ++i) SGG_TRY(refresh_shadow(theta, (__nv_bfloat16*)shadow, seg[i], (cudaStream_t)stream));
  return 0;
}

struct AdamHyper { float lr, b1, b2, eps; };
// Adam with the step number taken from the device counter: step = iter[0] * step_mul + step_add.
// which: 0 = every tensor, 1 = only W_a (the annotation rows of the attention kernel), 2 = everything but W_a
static int adam_dev(int net, const sgg_dims_t& d, float* theta, const float* grad, float* mm, float* vv, void* shadow,
                    const AdamHyper& h, const long long* iter, long long step_mul, long long step_add, cudaStream_t st,
                    int which = 0, long long wa_row0 = 0, long long wa_rows = -1) {
  const ParamLayout L = param_layout(net == 0, d);
  AdamParams p{};
  p.theta = theta; p.grad = grad; p.m = mm; p.v = vv; p.shadow = (__nv_bfloat16*)shadow;
  p.lr = h.lr; p.b1 = h.b1; p.b2 = h.b2; p.eps = h.eps; p.gscale = 1.0f;
  p.iter = iter; p.step_mul = step_mul; p.step_add = step_add;
  p.nseg = fill_adam_segs(L, p.seg);
  if (which == 1) {
    p.nseg = 1;                                   // segment 0 is W_a (fill_adam_segs)
    if (wa_rows >= 0) {                           // only the rows this rank owns (row-sharded projection)
      p.seg[0].off += wa_row0 * p.seg[0].cols;
      p.seg[0].sh_off += wa_row0 * p.seg[0].pitch;
      p.seg[0].n = wa_rows * p.seg[0].cols;
    }
  } else if (which == 2) {
    for (int i = 1; i < p.nseg; ++i) p.seg[i - 1] = p.seg[i];
    p.nseg -= 1;
  }
  return adam(p, st);
}

extern "C" int sgg_adam_step(int net, const sgg_dims_t* d, float* theta, const float* grad, float* m, float* v,
                             void* shadow, int64_t step, float lr, float beta1, float beta2, float eps,
                             float grad_scale, sgg_stream_t stream) {
  SGG_CHECK(d && theta && grad && m && v, "sgg_adam_step: null argument");
  SGG_CHECK(step >= 1, "sgg_adam_step: step must be >= 1");
  SGG_TRY(check_dims(*d));
  const ParamLayout L = param_layout(net == 0, *d);
  AdamParams p{};
  p.theta = theta; p.grad = grad; p.m = m; p.v = v; p.shadow = (__nv_bfloat16*)shadow;
  p.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step)));
  p.b1 = beta1; p.b2 = beta2; p.eps = eps; p.gscale = grad_scale;
  p.nseg = fill_adam_segs(L, p.seg);
  return adam(p, (cudaStream_t)stream);
}

// Op-level entry point of the optimiser-fused projection (adamproj.cu), see include/sgg_b200.h.
extern "C" int sgg_adam_project(int net, const sgg_dims_t* d, float* theta, const float* grad, float* m, float* v, void* shadow,
                                int64_t step, float lr, float beta1, float beta2, float eps, const void* ann, float* P,
                                int32_t write_shadow, sgg_stream_t stream) {
  SGG_CHECK(d && theta && grad && m && v && shadow && ann && P, "sgg_adam_project: null argument");
  SGG_CHECK(step >= 1, "sgg_adam_project: step must be >= 1");
  SGG_TRY(check_dims(*d));
  const ParamLayout L = param_layout(net == 0, *d);
  const Dm dm = derive(*d);
  SGG_CHECK(d->B <= 256 && (d->R & 3) == 0, "sgg_adam_project: needs B <= 256 and R %% 4 == 0 (got B=%d R=%d)", d->B, d->R);
  AdamProjParams ap{};
  ap.theta = theta + L.Watt; ap.grad = grad + L.Watt; ap.m = m + L.Watt; ap.v = v + L.Watt;
  ap.shadow = (__nv_bfloat16*)shadow + L.sWa; ap.pitch = L.pAtt; ap.lo_off = (long long)L.rWa * L.pAtt;
  ap.write_shadow = write_shadow ? 1 : 0;
  ap.lr = lr; ap.b1 = beta1; ap.b2 = beta2; ap.eps = eps;
  ap.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step)));
  ap.R = d->R; ap.total_kb = (int)((long long)d->R * d->C / 64); ap.M = d->B;
  ap.P = P; ap.ldP = dm.RP;
  SGG_TRY(zero_2d(P, dm.RP, d->R, d->B, (cudaStream_t)stream));
  return adam_proj(ap, ann, (cudaStream_t)stream);
}

extern "C" int sgg_rng_fill_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, sgg_stream_t stream) {
  SGG_CHECK(out || n == 0, "sgg_rng_fill_normal: null output");
  return rng_fill(out, n, seed, offset, 1, (cudaStream_t)stream);
}
extern "C" int sgg_rng_fill_uniform(float* out, int64_t n, uint64_t seed, uint64_t offset, sgg_stream_t stream) {
  SGG_CHECK(out || n == 0, "sgg_rng_fill_uniform: null output");
  return rng_fill(out, n, seed, offset, 0, (cudaStream_t)stream);
}

static int check_step(const sgg_step_args_t* a, bool need_d, bool need_labels, bool need_noise = true) {
  SGG_CHECK(a != nullptr, "step: null args");
  SGG_TRY(check_dims(a->dims));
  SGG_CHECK(a->workspace && a->workspace_bytes >= ws_layout(a->dims, nullptr).bytes,
            "step: workspace too small (%lld < %lld)", (long long)a->workspace_bytes,
            (long long)ws_layout(a->dims, nullptr).bytes);
  SGG_CHECK(a->g_theta && a->g_shadow && a->ann_g && (a->noise || !need_noise), "step: missing generator inputs");
  if (need_d) SGG_CHECK(a->d_theta && a->d_shadow && a->ann_d, "step: missing discriminator inputs");
  if (need_labels) SGG_CHECK(a->labels && (a->gp_alpha || !need_noise), "step: missing labels / gp_alpha");
  SGG_CHECK(a->world >= 1, "step: world must be >= 1");
  return 0;
}

// G forward only (gen:74-91 from self.downsampled): logits [B,T,V].
extern "C" int sgg_gen_forward(const sgg_step_args_t* a, sgg_stream_t stream) {
  SGG_TRY(check_step(a, false, false));
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = ws_layout(a->dims, a->workspace);
  const Net g = make_net(true, a->dims, a->g_theta, a->g_shadow, nullptr, a->ann_g, w.g, a->dims.B, st);
  return gen_forward(g, w, a->noise, 1, 0, (a->flags & SGG_FLAG_REFRESH_GEN_PROJ) != 0, a->logits_out);
}

// D forward on caller-supplied float triples [B,T,V] (disc:73-93): scores [B,T].
extern "C" int sgg_disc_forward(const sgg_step_args_t* a, const float* triples, float* scores_out, sgg_stream_t stream) {
  SGG_CHECK(a && triples && scores_out, "sgg_disc_forward: null argument");
  SGG_TRY(check_dims(a->dims));
  SGG_CHECK(a->d_theta && a->d_shadow && a->ann_d && a->workspace, "sgg_disc_forward: missing inputs");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = ws_layout(a->dims, a->workspace);
  const Net d = make_net(false, a->dims, a->d_theta, a->d_shadow, nullptr, a->ann_d, w.d, a->dims.B, st);
  const Dm& m = d.m;
  // triples [B,T,V] -> hi/lo rows t*B+b
  for (int t = 0; t < m.T; ++t) {
    PackParams pk{};
    pk.rows = m.B; pk.cols = m.V; pk.src = triples + (long long)t * m.V; pk.ld = (long long)m.T * m.V;
    pk.dst = w.TRIH + (long long)t * m.B * 2 * m.VP; pk.ldd = 2 * m.VP; pk.lo_off = m.VP;
    SGG_TRY(pack_hl(pk, st));
  }
  SGG_TRY(embed_dense(d, w, w.TRIH));
  EmbedMixParams em{};
  em.B = m.B; em.T = m.T; em.E = m.E; em.V = m.V; em.Uf = w.UF; em.ldUf = m.EP;
  em.blk_fake = 0; em.blk_real = -1; em.blk_int = -1;
  em.X = d.w.X; em.ldX = 2 * d.KXP; em.x_lo = d.KXP; em.strideT = d.sX(); em.uoff = d.uoff;
  SGG_TRY(embed_mix(em, st));
  SGG_TRY(net_attn_proj(d));
  SGG_TRY(net_init_state(d, 1));
  SGG_TRY(net_forward(d, 1));
  SGG_CUDA(cudaMemcpyAsync(scores_out, d.w.Y, (size_t)m.B * m.T * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// One discriminator step (train:365): grads of disc_cost = mean D(G(z)) - mean D(real) + lam * GP
// w.r.t. every Discriminator* variable into d_grad; scalars[1] = w_disc, scalars[2] = gp.
// `fake` = the generator's logits for this step (hi/lo rows t*B+b, a constant here), already computed.
// side_out: when non-null the dW_a chain is left running on the side stream returned through it (the caller enqueues
// the W_a optimiser work there and joins); when null the step joins before returning.
// pre: what the caller has already enqueued for this step's discriminator pass (sgg_train_iteration only):
//   proj_stream != null : P = flat(a_d) W_a is being computed on that stream (join it before the first scores GEMM)
//   state_ready         : c0 = h0 of the stream blocks is still in the workspace from the previous critic step
struct StepPre { cudaStream_t proj_stream; bool have_proj; bool state_ready; bool clear_P; };
static int disc_step_core(const sgg_step_args_t* a, const Ws& w, const __nv_bfloat16* fake, const float* gp_alpha,
                          float* scalars, cudaStream_t st, cudaStream_t* side_out = nullptr, const StepPre* pre = nullptr) {
  const sgg_dims_t& dd = a->dims;
  const Net d = make_net(false, dd, a->d_theta, a->d_shadow, a->d_grad, a->ann_d, w.d, 4 * dd.B, st, w.LNP);
  const Dm& m = d.m;
  const int B = m.B, T = m.T;
  const float invBT = 1.0f / ((float)B * a->world * T);
  bool pre_uf = false, pre_ea0 = false;
  {  // the dW_a block (first R*C*R floats of the bucket) is overwritten by its GEMM; everything else accumulates
    const long long skip = (long long)m.R * m.C * m.R;
    // one zero-fill launch for everything this step accumulates into before its first reverse pass: the gradient
    // bucket, the loss scalars, the embedding GEMM's output and the step-0 scores of the three streams
    ZeroList zl;
    if (zl.add(a->d_grad + skip, (d.L.total - skip) * 4) && zl.add(scalars, 16) &&
        zl.add(w.UF, (long long)T * B * m.EP * 4) && zl.add(d.w.EA, 3LL * B * m.RP * 4) &&
        zl.add(d.w.PB, (long long)B * m.RP * 4) &&   // P_bar: accumulated by the last reverse pass of this step
        // the tangent pass's start state (zero: c0 / h0 do not depend on the triples; the generator step uses the same
        // workspace with another row count), its embedding GEMM's output and its step-0 scores / gate rows
        zl.add(d.w.X + 3LL * B * 2 * d.KXP, (long long)B * 2 * d.KXP * 2) && zl.add(d.w.Cf + 3LL * B * m.H, (long long)B * m.H * 4) &&
        zl.add(d.w.CH + 3LL * B * 2 * m.H, (long long)B * 2 * m.H * 2) && zl.add(d.w.ED, (long long)B * m.RP * 4) &&
        zl.add(w.UF2, (long long)T * B * m.EP * 4) && zl.add(d.w.Q + 3LL * B * 4 * m.H, (long long)B * 4 * m.H * 4)) {
      SGG_TRY(zero_fill(zl, st));
      pre_uf = pre_ea0 = true;
    } else {
      SGG_CUDA(cudaMemsetAsync(a->d_grad + skip, 0, (size_t)(d.L.total - skip) * 4, st));
      SGG_CUDA(cudaMemsetAsync(scalars, 0, 4 * sizeof(float), st));
    }
  }
  // 2. embeddings of the three streams
  SGG_TRY(embed_dense(d, w, fake, pre_uf));
  EmbedMixParams em{};
  em.B = B; em.T = T; em.E = m.E; em.V = m.V; em.Uf = w.UF; em.ldUf = m.EP;
  em.labels = a->labels; em.Wemb = a->d_theta + d.L.Wemb; em.gp_alpha = gp_alpha;
  em.blk_fake = 0; em.blk_real = 1; em.blk_int = 2;
  em.X = d.w.X; em.ldX = 2 * d.KXP; em.x_lo = d.KXP; em.strideT = d.sX(); em.uoff = d.uoff;
  SGG_TRY(embed_mix(em, st));
  // 3. D forward on fake | real | interp (one annotation read per timestep)
  if (!(pre && pre->have_proj)) SGG_TRY(net_attn_proj(d));
  if (!(pre && pre->state_ready)) SGG_TRY(net_init_state(d, 3));
  if (pre && pre->have_proj && pre->proj_stream != st) SGG_TRY(side_join(st, pre->proj_stream));
  SGG_TRY(net_forward(d, 3, pre_ea0));
  LossParams lp{B, T, d.w.Y, 0, 1, invBT, scalars};
  {
    cudaStream_t ax = st;
    SGG_TRY(aux_fork(st, &ax));
    SGG_TRY(losses(lp, ax));
  }
  // The reverse pass of the fake and real streams is first order and needs nothing from the gradient-penalty chain
  // (interp data-path reverse -> slopes -> tangent forward -> reverse over interp + tangent): it runs on the auxiliary
  // stream beside that chain (both are sequences of small latency-bound kernels); the streams meet before the
  // weight-gradient GEMMs, which contract over all rows.  Measured (B200, config 2): 5.43 ms per iteration against 5.32 ms
  // with one reverse loop over all streams -- every kernel of either chain fills the SMs (one CTA per SM GEMMs, two-per-SM
  // attention), so the chains do not overlap and the split costs 60 launches and a second annotation read per step.
  // Hence off by default; SGG_TWO_STREAM_REV=1 enables it.
  static int two_env = -1;
  if (two_env < 0) { const char* e = getenv("SGG_TWO_STREAM_REV"); two_env = (e && e[0] == '1') ? 1 : 0; }
  cudaStream_t sB = st;
  RevCfg rfr{};
  rfr.blk0 = 0; rfr.nblk = 2; rfr.tan_pblk = -1; rfr.tan_blk = -1; rfr.ybar_blk[0] = invBT; rfr.ybar_blk[1] = -invBT; rfr.wgrad = true;
  const bool two = two_env != 0 && side_stream(1) != nullptr;
  rfr.ann_grad = a->ann_d_grad != nullptr;
  if (two) {
    SGG_TRY(zero_2d(d.w.PB, (long long)m.B * m.RP, (long long)m.B * m.RP, 1, st));   // both loops accumulate P_bar
    SGG_TRY(side_fork(st, &sB, 1));
    Net dB = d;
    dB.st = sB; dB.lnp = w.LNP2;
    SGG_TRY(net_reverse_loop(dB, rfr));
  }
  // 4. g = d sum D(x_hat) / d x_hat : data-path reverse on the interp block
  RevCfg ig{};
  ig.blk0 = 2; ig.nblk = 1; ig.tan_pblk = -1; ig.tan_blk = -1; ig.ybar_blk[2] = 1.0f; ig.wgrad = false;
  SGG_TRY(net_reverse(d, ig));
  SGG_TRY(embed_input_grad(d, w, 2, false));
  GpSlopesParams sp{B, T, m.V, w.DFAKE, m.VP, w.slopes, w.coef, scalars, 1.0f / ((float)B * a->world)};
  sp.vhl = w.VHL; sp.ldv = 2 * m.VP; sp.v_lo = m.VP;    // the tangent direction v = coef * g, packed hi/lo by the same kernel
  SGG_TRY(gp_slopes(sp, st));
  // 5. tangent forward along v = coef * g
  {
    bool pre_t = pre_uf;   // cleared by the step's first zero-fill launch (see above)
    if (!pre_t) {
      SGG_CUDA(cudaMemsetAsync(d.w.X + 3LL * B * 2 * d.KXP, 0, (size_t)B * 2 * d.KXP * 2, st));
      SGG_CUDA(cudaMemsetAsync(d.w.Cf + 3LL * B * m.H, 0, (size_t)B * m.H * 4, st));
      SGG_CUDA(cudaMemsetAsync(d.w.CH + 3LL * B * 2 * m.H, 0, (size_t)B * 2 * m.H * 2, st));
      SGG_CUDA(cudaMemsetAsync(d.w.ED, 0, (size_t)B * m.RP * 4, st));
    }
    SGG_TRY(embed_dense(d, w, w.VHL, pre_t, w.UF2));
    EmbedMixParams et{};
    et.B = B; et.T = T; et.E = m.E; et.V = m.V; et.Uf = w.UF2; et.ldUf = m.EP;
    et.blk_fake = 3; et.blk_real = -1; et.blk_int = -1;
    et.X = d.w.X; et.ldX = 2 * d.KXP; et.x_lo = d.KXP; et.strideT = d.sX(); et.uoff = d.uoff;
    SGG_TRY(embed_mix(et, st));
    SGG_TRY(net_tangent(d, 2, 3, pre_t));
  }
  // 6. reverse over the primal streams and the tangent
  cudaStream_t s1 = st;
  if (two) {
    RevCfg rit{};
    rit.blk0 = 2; rit.nblk = 1; rit.tan_pblk = 2; rit.tan_blk = 3; rit.ybar_blk[2] = 0.f; rit.ydot_bar = a->lam; rit.wgrad = true;
    rit.ann_grad = a->ann_d_grad != nullptr;
    SGG_TRY(net_reverse_loop(d, rit));
    SGG_TRY(side_join(st, sB, 1));
    SGG_TRY(net_reverse_wgrad(d, pre && pre->clear_P, &s1, w.LNP, rev_slices(d, rit), w.LNP2, rev_slices(d, rfr)));
  } else {
    RevCfg rv{};
    rv.blk0 = 0; rv.nblk = 3; rv.tan_pblk = 2; rv.tan_blk = 3;
    rv.ybar_blk[0] = invBT; rv.ybar_blk[1] = -invBT; rv.ybar_blk[2] = 0.f; rv.ydot_bar = a->lam;
    rv.wgrad = true;
    rv.pb_prezeroed = pre_uf;
    rv.clear_P = pre && pre->clear_P;
    rv.ann_grad = a->ann_d_grad != nullptr;
    SGG_TRY(net_reverse(d, rv, &s1));
  }
  // 7. embedding gradient: fake^T (ub_f + al ub_i) + scatter(labels, ub_r + (1-al) ub_i) + v^T udot_bar
  {
    PackParams pk{};
    pk.rows = T * B; pk.cols = m.E; pk.rpg = B;
    pk.src = d.w.XB + d.uoff; pk.ld = d.KXP; pk.gstride = d.sXB();
    pk.src2 = d.w.XB + 2LL * B * d.KXP + d.uoff; pk.ld2 = d.KXP; pk.gstride2 = d.sXB();
    pk.mix = gp_alpha; pk.mmod = B;
    pk.dst = w.UBH; pk.ldd = 2 * m.EP; pk.lo_off = m.EP;
    SGG_TRY(pack_hl(pk, st));
    PackParams pt{};
    pt.rows = T * B; pt.cols = m.E; pt.rpg = B;
    pt.src = d.w.XB + 3LL * B * d.KXP + d.uoff; pt.ld = d.KXP; pt.gstride = d.sXB();
    pt.dst = w.UDB; pt.ldd = 2 * m.EP; pt.lo_off = m.EP;
    SGG_TRY(pack_hl(pt, st));
    for (int which = 0; which < 2; ++which) {
      sgg_gemm_desc_t q = gd_zero();
      q.A = which == 0 ? fake : w.VHL; q.a_rows = (long long)T * B; q.a_cols = 2 * m.VP; q.a_ld = 2 * m.VP; q.a_mn_major = 1;
      q.B = which == 0 ? w.UBH : w.UDB; q.b_rows = (long long)T * B; q.b_cols = 2 * m.EP; q.b_ld = 2 * m.EP; q.b_mn_major = 1;
      q.M = m.V; q.N = m.E; q.nseg = 3;
      for (int s = 0; s < 3; ++s) q.seg_klen[s] = T * B;
      q.seg_b_mn[1] = m.EP; q.seg_a_mn[2] = m.VP;
      q.C = a->d_grad + d.L.Wemb; q.ldc = m.E; q.atomic = 1;
      SGG_TRY(gemm(q, st));
    }
    EmbedScatterParams es{};
    es.B = B; es.T = T; es.E = m.E; es.V = m.V; es.labels = a->labels; es.gp_alpha = gp_alpha;
    es.XB = d.w.XB; es.ldXB = d.KXP; es.strideT = d.sXB(); es.uoff = d.uoff;
    es.blk_real = 1; es.blk_int = 2; es.dWemb = a->d_grad + d.L.Wemb;
    cudaStream_t ax = st;
    SGG_TRY(aux_fork(st, &ax));          // the x_bar rows it reads are complete; its reductions commute with the GEMMs'
    SGG_TRY(embed_scatter(es, ax));
  }
  SGG_TRY(aux_join(st));
  if (a->ann_d_grad) {   // d disc_cost / d ann_d for the caller's conv front-end (disc:29-68); step-level entry point only
    SGG_CHECK(side_out == nullptr, "annotation gradients are produced by sgg_disc_step / sgg_gen_step, not by sgg_train_iteration");
    SGG_TRY(side_join(st, s1));
    return net_ann_grad(d, 3, 3, a->ann_d_grad);
  }
  if (side_out) *side_out = s1;
  else SGG_TRY(side_join(st, s1));
  return 0;
}

extern "C" int sgg_disc_step(const sgg_step_args_t* a, sgg_stream_t stream) {
  SGG_TRY(check_step(a, true, true));
  SGG_CHECK(a->d_grad && a->scalars, "sgg_disc_step: missing d_grad / scalars");
  cudaStream_t st = (cudaStream_t)stream;
  const sgg_dims_t& dd = a->dims;
  const Ws w = ws_layout(dd, a->workspace);
  const Net g = make_net(true, dd, a->g_theta, a->g_shadow, nullptr, a->ann_g, w.g, dd.B, st);
  // 1. fake = G(a_g, noise)  (constant for this step)
  SGG_TRY(gen_forward(g, w, a->noise, 1, 0, (a->flags & SGG_FLAG_REFRESH_GEN_PROJ) != 0, a->logits_out));
  return disc_step_core(a, w, fake_slot(w, g.m, 0), a->gp_alpha, a->scalars, st);
}

// One generator step (train:368): grads of gen_cost = -mean D(G(z)) w.r.t. every Generator*
// variable into g_grad; scalars[3] = gen_cost.
static int gen_step_core(const sgg_step_args_t* a, const Ws& w, const float* noise, bool recompute_proj, float* scalars,
                         cudaStream_t st, cudaStream_t* side_out = nullptr, const StepPre* pre = nullptr) {
  const sgg_dims_t& dd = a->dims;
  const Net g = make_net(true, dd, a->g_theta, a->g_shadow, a->g_grad, a->ann_g, w.g, dd.B, st, w.LNP);
  const Net d = make_net(false, dd, a->d_theta, a->d_shadow, nullptr, a->ann_d, w.d, dd.B, st);
  const Dm& m = d.m;
  const int B = m.B, T = m.T;
  const float invBT = 1.0f / ((float)B * a->world * T);
  bool g_pb_zeroed = false;
  {
    const long long skip = (long long)m.R * m.C * m.R;   // dW_a is overwritten by its GEMM
    ZeroList zl;
    if (zl.add(a->g_grad + skip, (g.L.total - skip) * 4) && zl.add(scalars, 16) && zl.add(g.w.PB, (long long)B * m.RP * 4)) {
      SGG_TRY(zero_fill(zl, st));
      g_pb_zeroed = true;
    } else {
      SGG_CUDA(cudaMemsetAsync(a->g_grad + skip, 0, (size_t)(g.L.total - skip) * 4, st));
      SGG_CUDA(cudaMemsetAsync(scalars, 0, 4 * sizeof(float), st));
    }
  }
  SGG_TRY(gen_forward(g, w, noise, 1, 0, recompute_proj, a->logits_out));
  // D(fake), single stream
  SGG_TRY(embed_dense(d, w, fake_slot(w, m, 0)));
  EmbedMixParams em{};
  em.B = B; em.T = T; em.E = m.E; em.V = m.V; em.Uf = w.UF; em.ldUf = m.EP;
  em.blk_fake = 0; em.blk_real = -1; em.blk_int = -1;
  em.X = d.w.X; em.ldX = 2 * d.KXP; em.x_lo = d.KXP; em.strideT = d.sX(); em.uoff = d.uoff;
  SGG_TRY(embed_mix(em, st));
  if (!(pre && pre->have_proj)) SGG_TRY(net_attn_proj(d));
  SGG_TRY(net_init_state(d, 1));
  if (pre && pre->have_proj && pre->proj_stream != st) SGG_TRY(side_join(st, pre->proj_stream));
  SGG_TRY(net_forward(d, 1));
  LossParams lp{B, T, d.w.Y, 0, -1, invBT, scalars};
  {
    cudaStream_t ax = st;
    SGG_TRY(aux_fork(st, &ax));
    SGG_TRY(losses(lp, ax));
  }
  // d gen_cost / d fake through D's data path
  RevCfg rd{};
  rd.blk0 = 0; rd.nblk = 1; rd.tan_pblk = -1; rd.tan_blk = -1; rd.ybar_blk[0] = -invBT; rd.wgrad = false;
  SGG_TRY(net_reverse(d, rd));
  SGG_TRY(embed_input_grad(d, w, 0, true));
  // G reverse: h_bar from the logits for all t in one GEMM, then the time loop, then weight grads
  {
    sgg_gemm_desc_t q = gd_zero();
    q.A = w.DFAKEH; q.a_rows = (long long)T * B; q.a_cols = 2 * m.VP; q.a_ld = 2 * m.VP;
    q.B = g.sh + g.L.sWdec; q.b_rows = 2LL * g.L.rWdec; q.b_cols = m.V; q.b_ld = g.L.pWdec; q.b_mn_major = 0;
    q.M = T * B; q.N = m.H;
    segs_act_weight(q, 0, m.VP, g.L.rWdec, false, m.VP);
    q.C = w.HB; q.ldc = m.H;
    SGG_TRY(gemm(q, st));
  }
  RevCfg rg{};
  rg.blk0 = 0; rg.nblk = 1; rg.tan_pblk = -1; rg.tan_blk = -1; rg.HB = w.HB; rg.wgrad = true; rg.pb_prezeroed = g_pb_zeroed;
  rg.ann_grad = a->ann_g_grad != nullptr;
  cudaStream_t s1 = st;
  SGG_TRY(net_reverse(g, rg, &s1));
  {  // dW_dec = H^T dfake [H, V], db_dec = column sums of dfake
    sgg_gemm_desc_t q = gd_zero();
    q.A = g.w.X + g.sX(); q.a_rows = (long long)T * B; q.a_cols = 2 * g.KXP; q.a_ld = 2 * g.KXP; q.a_mn_major = 1;
    q.B = w.DFAKEH; q.b_rows = (long long)T * B; q.b_cols = 2 * m.VP; q.b_ld = 2 * m.VP; q.b_mn_major = 1;
    q.M = m.H; q.N = m.V; q.nseg = 3;
    for (int s = 0; s < 3; ++s) { q.seg_klen[s] = T * B; q.seg_a_mn[s] = g.hoff; }
    q.seg_b_mn[1] = m.VP; q.seg_a_mn[2] = g.KXP + g.hoff;
    q.C = a->g_grad + g.L.Wdec; q.ldc = m.V; q.atomic = 1;
    SGG_TRY(gemm(q, st));
    cudaStream_t ax = st;
    SGG_TRY(aux_fork(st, &ax));
    SGG_TRY(colsum(w.DFAKE, m.VP, T * B, m.V, a->g_grad + g.L.bdec, ax));
  }
  SGG_TRY(aux_join(st));
  if (a->ann_g_grad) {   // d gen_cost / d ann_g for the caller's conv front-end (gen:29-68); step-level entry point only
    SGG_CHECK(side_out == nullptr, "annotation gradients are produced by sgg_disc_step / sgg_gen_step, not by sgg_train_iteration");
    SGG_TRY(side_join(st, s1));
    return net_ann_grad(g, 1, -1, a->ann_g_grad);
  }
  if (side_out) *side_out = s1;
  else SGG_TRY(side_join(st, s1));
  return 0;
}

extern "C" int sgg_gen_step(const sgg_step_args_t* a, sgg_stream_t stream) {
  SGG_TRY(check_step(a, true, false));
  SGG_CHECK(a->g_grad && a->scalars, "sgg_gen_step: missing g_grad / scalars");
  const Ws w = ws_layout(a->dims, a->workspace);
  return gen_step_core(a, w, a->noise, (a->flags & SGG_FLAG_REFRESH_GEN_PROJ) != 0, a->scalars, (cudaStream_t)stream);
}

// ============================================================================ row-sharded projection: buffers
namespace sgg { int slab_pack(const __nv_bfloat16* a, __nv_bfloat16* send, int B, long long K, long long Ks, int world, cudaStream_t st); }
struct ShardScratch { __nv_bfloat16* send; float* Ppart[2]; float* PBall[2]; __nv_bfloat16* PBHall[2]; long long bytes; };
static long long shard_ks(const sgg_dims_t& d, int world) {
  const long long kb = ((long long)d.R * d.C + 63) / 64;
  return (kb + world - 1) / world * 64;
}
static ShardScratch shard_scratch_layout(const sgg_dims_t& d, int world, void* base) {
  const Dm m = derive(d);
  const long long Ks = shard_ks(d, world), Bg = (long long)m.B * world;
  ShardScratch s{};
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  long long o = 0;
  auto take = [&](long long bytes) { uint8_t* r = p ? p + o : nullptr; o = rup(o + bytes, 256); return (void*)r; };
  s.send = (__nv_bfloat16*)take(2LL * world * m.B * Ks * 2);   // generator and discriminator staging
  for (int i = 0; i < 2; ++i) {
    s.Ppart[i] = (float*)take(Bg * m.RP * 4);
    s.PBall[i] = (float*)take(Bg * m.RP * 4);
    s.PBHall[i] = (__nv_bfloat16*)take(Bg * 2 * m.RP * 2);
  }
  s.bytes = o;
  return s;
}
static int shard_ctx(const sgg_dims_t& d, int world, int rank, void* comm, const sgg_wa_shard_t& sh, ShardCtx* out) {
  SGG_CHECK(((long long)d.R * d.C / 64) % world == 0, "row-sharded projection: R*C/64 = %lld is not divisible by world = %d",
            (long long)d.R * d.C / 64, world);
  SGG_CHECK(sh.slab_g && sh.slab_d && sh.scratch, "row-sharded projection: missing slab / scratch buffers");
  const ShardScratch s = shard_scratch_layout(d, world, sh.scratch);
  SGG_CHECK(sh.scratch_bytes >= s.bytes, "row-sharded projection: scratch too small (%lld < %lld)", (long long)sh.scratch_bytes,
            (long long)s.bytes);
  out->comm = comm; out->rank = rank; out->world = world; out->Ks = shard_ks(d, world);
  out->slab_g = (const __nv_bfloat16*)sh.slab_g; out->slab_d = (const __nv_bfloat16*)sh.slab_d;
  out->send = s.send;
  for (int i = 0; i < 2; ++i) { out->Ppart[i] = s.Ppart[i]; out->PBall[i] = s.PBall[i]; out->PBHall[i] = s.PBHall[i]; }
  return 0;
}
extern "C" int64_t sgg_wa_shard_scratch_bytes(const sgg_dims_t* d, int32_t world) {
  if (!d || check_dims(*d) != 0 || world < 1) return -1;
  return shard_scratch_layout(*d, world, nullptr).bytes;
}
extern "C" int64_t sgg_wa_shard_slab_elems(const sgg_dims_t* d, int32_t world) {
  if (!d || check_dims(*d) != 0 || world < 1) return -1;
  return (int64_t)d->B * world * shard_ks(*d, world);
}

// ============================================================================ one training iteration
// train:362-368 on one batch (train:185-187): critic_iters x {D step, Adam(D)} then {G step, Adam(G)}, with
// fresh noise / interpolation coefficients per step drawn on the device.  The generator is constant during the
// critic steps, so its forwards for ALL critic steps run first as one batched forward (their noise draws
// share each annotation tile read).  Everything is enqueued on `stream` and depends on the host only through
// the arguments, so one call can be captured into a CUDA graph and replayed: the iteration counter, RNG
// position and Adam step numbers live in device memory (`counters`).
extern "C" int sgg_train_iteration(const sgg_iter_args_t* it, sgg_stream_t stream) {
  SGG_CHECK(it != nullptr, "sgg_train_iteration: null args");
  const sgg_step_args_t* a = &it->step;
  SGG_TRY(check_step(a, true, true, false));
  SGG_CHECK(a->g_grad && a->d_grad, "sgg_train_iteration: missing gradient buckets");
  SGG_CHECK(it->g_m && it->g_v && it->d_m && it->d_v, "sgg_train_iteration: missing Adam moments");
  SGG_CHECK(it->critic_iters >= 0 && it->critic_iters <= a->dims.S, "sgg_train_iteration: critic_iters=%d exceeds dims.S=%d",
            it->critic_iters, a->dims.S);
  SGG_CHECK(it->counters && it->noise_all && it->gp_alpha_all && it->scalars_all, "sgg_train_iteration: missing buffers");
  SGG_CHECK(!a->ann_g_grad && !a->ann_d_grad, "sgg_train_iteration: annotation gradients are produced by sgg_disc_step / sgg_gen_step only");
  cudaStream_t st = (cudaStream_t)stream;
  AnnStaticScope ann_static;
  const sgg_dims_t& dd = a->dims;
  const Ws w = ws_layout(dd, a->workspace);
  const Dm m = derive(dd);
  const int nc = it->critic_iters;
  const long long* iter = reinterpret_cast<const long long*>(it->counters);
  // ---- row-sharded attention projection (world > 1): exchange the annotation column slabs of this batch
  ShardCtx sc{};
  const bool shard = it->comm && a->world > 1 && it->shard.enabled;
  if (shard) {
    SGG_CHECK(comm_world(it->comm) == a->world, "sgg_train_iteration: communicator has %d ranks, step.world = %d",
              comm_world(it->comm), a->world);
    SGG_TRY(shard_ctx(dd, a->world, comm_rank(it->comm), it->comm, it->shard, &sc));
  }
  ShardScope shard_scope(shard ? &sc : nullptr);
  cudaStream_t s_ex = st;
  if (shard) {
    const long long per = (long long)m.B * sc.Ks;
    SGG_TRY(slab_pack((const __nv_bfloat16*)a->ann_g, sc.send, m.B, (long long)m.R * m.C, sc.Ks, a->world, st));
    SGG_TRY(comm_all_to_all(it->comm, sc.send, const_cast<__nv_bfloat16*>(sc.slab_g), per * 2, st, 0));
    // the discriminator's slab is first needed by the first critic step: exchange it beside the generator forwards
    if (comm_has_side_lane(it->comm)) SGG_TRY(side_fork(st, &s_ex));
    __nv_bfloat16* send_d = sc.send + (long long)a->world * per;
    SGG_TRY(slab_pack((const __nv_bfloat16*)a->ann_d, send_d, m.B, (long long)m.R * m.C, sc.Ks, a->world, s_ex));
    SGG_TRY(comm_all_to_all(it->comm, send_d, const_cast<__nv_bfloat16*>(sc.slab_d), per * 2, s_ex, s_ex != st ? 1 : 0));
  }
  // ---- first critic step's projection / initial state on the side stream, beside the generator forwards
  static int prefetch_proj = -1;             // SGG_PROJ_PREFETCH=0 disables these overlaps (A/B measurements)
  if (prefetch_proj < 0) { const char* e = getenv("SGG_PROJ_PREFETCH"); prefetch_proj = (e && e[0] == '0') ? 0 : 1; }
  const bool side_ok = !it->comm || comm_has_side_lane(it->comm);
  const bool use_side = !shard || side_ok;   // sharded: the side chain issues collectives, which need their own lane
  cudaStream_t s_pre = s_ex;
  if (prefetch_proj != 0 && nc > 0 && use_side && side_stream() != nullptr) {
    if (s_pre == st) SGG_TRY(side_fork(st, &s_pre));
    const Net dn = make_net(false, dd, a->d_theta, a->d_shadow, nullptr, a->ann_d, w.d, 4 * m.B, s_pre);
    SGG_TRY(net_attn_proj(dn, s_pre != st ? 1 : 0));
    SGG_TRY(net_init_state(dn, 3));
  }
  // ---- fresh randomness for every step of this iteration (Philox position = f(iteration counter))
  const long long n_noise = (long long)(nc + 1) * m.B * m.C, n_alpha = (long long)nc * m.B;
  const uint64_t per_iter = (uint64_t)((n_noise + 3) / 4 + (n_alpha + 3) / 4);
  SGG_TRY(rng_fill(it->noise_all, n_noise, it->seed, 0, 1, st, iter, per_iter));
  SGG_TRY(rng_fill(it->gp_alpha_all, n_alpha, it->seed, (uint64_t)((n_noise + 3) / 4), 0, st, iter, per_iter));
  // ---- generator forwards of all critic steps, MAX_GEN_STREAMS draws per pass
  bool fresh = true;
  for (int s0 = 0; s0 < nc; s0 += m.GS) {
    const int ns = nc - s0 < m.GS ? nc - s0 : m.GS;
    const Net g = make_net(true, dd, a->g_theta, a->g_shadow, nullptr, a->ann_g, w.g, ns * m.B, st);
    SGG_TRY(gen_forward(g, w, it->noise_all + (long long)s0 * m.B * m.C, ns, s0, fresh, nullptr));
    fresh = false;
  }
  // ---- critic steps
  AdamHyper hp{it->lr, it->beta1, it->beta2, it->eps};
  // Each optimiser step is split in two independent chains that meet again before the next step:
  //   side stream : dW_a GEMM -> [all-reduce of the W_a block] -> Adam on W_a           (HBM- / NVLink-bound)
  //   main stream : the other weight gradients -> [all-reduce of the rest] -> Adam on the rest
  // `next_proj`: the discriminator's projection for the NEXT pass (P = flat(a_d) W_a with the updated W_a) is enqueued
  // right behind the W_a update on the side stream, where it overlaps the main stream's weight-gradient GEMMs; the
  // side stream is then joined by the next pass just before its first scores GEMM instead of here.
  static int fuse_env = -1;   // SGG_FUSED_ADAM_PROJ=0 keeps the Adam update of W_a and the next projection as two kernels
  if (fuse_env < 0) { const char* e = getenv("SGG_FUSED_ADAM_PROJ"); fuse_env = (e && e[0] == '0') ? 0 : 1; }
  const bool will_fuse = fuse_env && prefetch_proj != 0 && !shard && !it->comm && m.B <= 256 && (m.R & 3) == 0 &&
                         ((long long)m.R * m.C) % 64 == 0;
  auto optimise = [&](int net, float* theta, float* grad, float* mm, float* vv, void* shadow, long long step_mul,
                      long long step_add, cudaStream_t s1, bool next_proj, cudaStream_t* pending, bool last_critic) -> int {
    const ParamLayout L = param_layout(net == 0, dd);
    const long long n_wa = (long long)m.R * m.C * m.R;
    if (s1 != st && !side_ok) { SGG_TRY(side_join(st, s1)); s1 = st; }
    bool fused = false;
    if (will_fuse && next_proj && net == 1) {
      // Adam on W_a and the next pass's P = flat(a_d) W_a in one kernel (adamproj.cu): the updated weights go from
      // registers into the MMA operand; the hi/lo shadow is only written for the pass that will see new annotations
      AdamProjParams ap{};
      ap.theta = theta + L.Watt; ap.grad = grad + L.Watt; ap.m = mm + L.Watt; ap.v = vv + L.Watt;
      ap.shadow = (__nv_bfloat16*)shadow + L.sWa; ap.pitch = L.pAtt; ap.lo_off = (long long)L.rWa * L.pAtt;
      ap.write_shadow = last_critic ? 1 : 0;
      ap.lr = hp.lr; ap.b1 = hp.b1; ap.b2 = hp.b2; ap.eps = hp.eps;
      ap.iter = iter; ap.step_mul = step_mul; ap.step_add = step_add;
      ap.R = m.R; ap.total_kb = (int)((long long)m.R * m.C / 64); ap.M = m.B;
      ap.P = w.d.P; ap.ldP = m.RP;
      SGG_TRY(adam_proj(ap, a->ann_d, s1));   // P was cleared by the W_a chain's packing kernel (RevCfg::clear_P)
      fused = true;
    } else if (shard) {   // this rank's rows of dW_a are already sums over the global batch: no exchange, 1/world of the update
      SGG_TRY(adam_dev(net, dd, theta, grad, mm, vv, shadow, hp, iter, step_mul, step_add, s1, 1, sc.rank * sc.Ks, sc.Ks));
    } else {
      if (it->comm) SGG_TRY(comm_allreduce(it->comm, grad, n_wa, s1, s1 != st ? 1 : 0));
      SGG_TRY(adam_dev(net, dd, theta, grad, mm, vv, shadow, hp, iter, step_mul, step_add, s1, 1));
    }
    if (next_proj && !fused) {
      const Net dn = make_net(false, dd, a->d_theta, a->d_shadow, nullptr, a->ann_d, w.d, m.B, s1);
      SGG_TRY(net_attn_proj(dn, s1 != st ? 1 : 0));
    }
    if (it->comm) SGG_TRY(comm_allreduce(it->comm, grad + n_wa, L.total - n_wa, st, 0));
    SGG_TRY(adam_dev(net, dd, theta, grad, mm, vv, shadow, hp, iter, step_mul, step_add, st, 2));
    if (next_proj && pending) { *pending = s1; return 0; }
    return side_join(st, s1);
  };
  StepPre pre{st, false, false, will_fuse};
  if (prefetch_proj != 0 && nc > 0 && use_side && side_stream() != nullptr) {
    // the first critic step's projection and initial state do not depend on the generator: they were enqueued on the
    // side stream (behind the discriminator's slab exchange when sharded) and ran beside the generator forwards
    pre.proj_stream = s_pre; pre.have_proj = true; pre.state_ready = true;
  } else if (s_ex != st) {
    SGG_TRY(side_join(st, s_ex));
  }
  for (int i = 0; i < nc; ++i) {
    cudaStream_t s1 = st;
    SGG_TRY(disc_step_core(a, w, fake_slot(w, m, i), it->gp_alpha_all + (long long)i * m.B, it->scalars_all + 4 * i, st,
                           use_side ? &s1 : nullptr, &pre));
    cudaStream_t pending = st;
    SGG_TRY(optimise(1, const_cast<float*>(a->d_theta), a->d_grad, it->d_m, it->d_v, const_cast<void*>(a->d_shadow), nc, i + 1, s1,
                     prefetch_proj != 0, &pending, i == nc - 1));
    pre.have_proj = prefetch_proj != 0; pre.proj_stream = pending;
    pre.state_ready = true;                  // same batch, same row layout: c0 = h0 stay valid for the next critic step
  }
  // ---- generator step (its own forward keeps the activations the reverse pass needs)
  {
    cudaStream_t s1 = st;
    SGG_TRY(gen_step_core(a, w, it->noise_all + (long long)nc * m.B * m.C, fresh, it->scalars_all + 4 * nc, st,
                          use_side ? &s1 : nullptr, &pre));
    SGG_TRY(optimise(0, const_cast<float*>(a->g_theta), a->g_grad, it->g_m, it->g_v, const_cast<void*>(a->g_shadow), 1, 1, s1,
                     false, nullptr, false));
  }
  return bump_counter(reinterpret_cast<long long*>(it->counters), st);
}

// ============================================================================ generator sampling (inference)
// Forward-only generator (gen:74-91) followed by the reference's test-time decoding (train:270 tf.argmax over the
// vocabulary) or Gumbel-max sampling.  Nothing is kept for a reverse pass: the batch is walked in chunks whose state,
// input and scratch buffers are reused every timestep, the hoisted projection P is computed once for the whole batch
// and the vocabulary logits never reach HBM unless the caller asks for them -- the decoder GEMM's epilogue reduces each
// n-tile to a (max, column) key.
struct SampleWs {
  float* P; float* noise; unsigned long long* keys;
  NetWs g[2];      // two chunk workspaces: consecutive chunks alternate between the caller's stream and the side stream
  long long bytes;
};
// Default: two chunks per call (B >= 2048), so that the HBM-bound attention steps of one half overlap the tensor-bound
// GEMMs of the other half on the second stream; small batches run as one chunk.
static int sample_chunk(const sgg_dims_t& d, int chunk) {
  int c = chunk > 0 ? chunk : (d.B >= 2048 ? (d.B + 1) / 2 : d.B);
  if (c > 8192) c = 8192;
  return c < d.B ? c : d.B;
}
static SampleWs sample_ws_layout(const sgg_dims_t& d, int chunk, void* base) {
  const Dm m = derive(d);
  const int Bc = sample_chunk(d, chunk);
  SampleWs w{};
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  long long o = 0;
  auto take = [&](long long bytes) { uint8_t* r = p ? p + o : nullptr; o = rup(o + bytes, 256); return (void*)r; };
  w.P = (float*)take((long long)m.B * m.RP * 4);
  w.noise = (float*)take((long long)m.B * m.C * 4);
  w.keys = (unsigned long long*)take((long long)m.B * m.T * 8);
  for (int i = 0; i < (Bc < m.B ? 2 : 1); ++i) {
    w.g[i].NRmax = Bc;
    w.g[i].X = (__nv_bfloat16*)take(2LL * Bc * 2 * m.KXG * 2);
    w.g[i].Cf = (float*)take(2LL * Bc * m.H * 4);
    w.g[i].CH = (__nv_bfloat16*)take(2LL * Bc * 2 * m.H * 2);
    w.g[i].EA = (float*)take((long long)Bc * m.RP * 4);
    w.g[i].Q = (float*)take((long long)Bc * 4 * m.H * 4);
    w.g[i].GSCR = (float*)take(gates_scratch_floats(Bc) * 4);
  }
  w.bytes = o;
  return w;
}

namespace sgg { int decode_keys(const unsigned long long* keys, int32_t* tokens, long long n, cudaStream_t st); }

extern "C" int64_t sgg_sample_workspace_bytes(const sgg_dims_t* d, int32_t chunk) {
  if (!d || check_dims(*d) != 0) return -1;
  return sample_ws_layout(*d, chunk, nullptr).bytes;
}

extern "C" int sgg_gen_sample(const sgg_sample_args_t* a, sgg_stream_t stream) {
  SGG_CHECK(a != nullptr, "sgg_gen_sample: null args");
  SGG_TRY(check_dims(a->dims));
  SGG_CHECK(a->g_theta && a->g_shadow && a->ann_g && a->tokens_out, "sgg_gen_sample: missing inputs / outputs");
  SGG_CHECK(a->mode == SGG_SAMPLE_GREEDY || a->mode == SGG_SAMPLE_GUMBEL, "sgg_gen_sample: unknown mode %d", a->mode);
  cudaStream_t st = (cudaStream_t)stream;
  const sgg_dims_t& dd = a->dims;
  const SampleWs w = sample_ws_layout(dd, a->chunk, a->workspace);
  SGG_CHECK(a->workspace && a->workspace_bytes >= w.bytes, "sgg_gen_sample: workspace too small (%lld < %lld)",
            (long long)a->workspace_bytes, (long long)w.bytes);
  const Dm m = derive(dd);
  const int Bc = sample_chunk(dd, a->chunk);
  AnnStaticScope ann_static;   // the annotations are an input of the whole call
  const float* noise = a->noise;
  if (!noise) {  // gen:81: one N(0,1) draw per image, shared by all timesteps
    SGG_TRY(rng_fill(w.noise, (long long)m.B * m.C, a->seed, a->offset, 1, st));
    noise = w.noise;
  }
  SGG_CUDA(cudaMemsetAsync(w.keys, 0, (size_t)m.B * m.T * 8, st));
  {  // K1 for the whole batch: W_a streams through the SMs once (tensor-bound at large B)
    NetWs wa = w.g[0]; wa.P = w.P;
    const Net g = make_net(true, dd, a->g_theta, a->g_shadow, nullptr, a->ann_g, wa, m.B, st);
    SGG_TRY(net_attn_proj(g));
  }
  static int two_streams = -1;   // SGG_SAMPLE_STREAMS=1: every chunk on the caller's stream
  if (two_streams < 0) { const char* e = getenv("SGG_SAMPLE_STREAMS"); two_streams = (e && e[0] == '1') ? 0 : 1; }
  cudaStream_t s2 = st;
  if (two_streams && Bc < m.B) SGG_TRY(side_fork(st, &s2));
  int ci = 0;
  for (int c0 = 0; c0 < m.B; c0 += Bc, ++ci) {
    cudaStream_t sc = (ci & 1) ? s2 : st;
    sgg_dims_t dc = dd;
    dc.B = m.B - c0 < Bc ? m.B - c0 : Bc;
    dc.S = 1;
    NetWs wc = w.g[(Bc < m.B) ? (ci & 1) : 0];
    wc.P = w.P + (long long)c0 * m.RP;
    Net g = make_net(true, dc, a->g_theta, a->g_shadow, nullptr,
                     (const __nv_bfloat16*)a->ann_g + (long long)c0 * m.R * m.C, wc, dc.B, sc);
    g.roll = true;
    SGG_TRY(net_init_state(g, 1));
    {
      PackParams pk{};
      pk.rows = dc.B; pk.cols = m.C; pk.src = noise + (long long)c0 * m.C; pk.ld = m.C;
      pk.dst = g.w.X + g.uoff; pk.ldd = 2 * g.KXP; pk.lo_off = g.KXP;
      pk.reps = 2; pk.rep_stride = g.sX();
      SGG_TRY(pack_hl(pk, sc));
    }
    for (int t = 0; t < m.T; ++t) {
      SGG_TRY(net_forward_step(g, t, 1));
      // logits_t = h_{t+1} W_dec + b (gen:88) with the decoding fused into the epilogue
      sgg_gemm_desc_t q = gd_zero();
      q.A = g.w.X + g.t2(t + 1) * g.sX(); q.a_rows = dc.B; q.a_cols = 2 * g.KXP; q.a_ld = 2 * g.KXP;
      q.B = g.sh + g.L.sWdec; q.b_rows = 2LL * g.L.rWdec; q.b_cols = m.V; q.b_ld = g.L.pWdec; q.b_mn_major = 1;
      q.M = dc.B; q.N = m.V;
      segs_act_weight(q, g.hoff, g.KXP, g.L.rWdec, true, m.H);
      q.bias = g.theta + g.L.bdec;
      q.splits = 1;
      q.argmax_keys = reinterpret_cast<uint64_t*>(w.keys + (long long)c0 * m.T + t); q.argmax_stride = m.T;
      q.gumbel = a->mode == SGG_SAMPLE_GUMBEL; q.gumbel_seed = a->seed ^ 0x5851F42D4C957F2DULL;
      q.gumbel_offset = a->offset + (unsigned long long)c0 + (unsigned long long)t * (unsigned long long)m.B;
      if (a->logits_out) { q.C = a->logits_out + ((long long)c0 * m.T + t) * m.V; q.ldc = (long long)m.T * m.V; }
      SGG_TRY(gemm(q, sc));
    }
  }
  SGG_TRY(side_join(st, s2));
  return decode_keys(w.keys, a->tokens_out, (long long)m.B * m.T, st);
}

// Run-time switches (same meaning as the environment variables read at first use; for A/B tests inside one process).
extern "C" int sgg_set_option(const char* name, int32_t value) {
  SGG_CHECK(name != nullptr, "sgg_set_option: null name");
  if (strcmp(name, "fused_gates") == 0) { gates_set_mode(value); return 0; }
  if (strcmp(name, "attn_persistent") == 0) { attn_set_persistent(value); return 0; }
  set_error("sgg_set_option: unknown option '%s'", name);
  return -1;
}

// Debug / test accessor: location of an intermediate buffer inside the workspace.
extern "C" int sgg_ws_lookup(const sgg_dims_t* d, const char* name, int64_t* offset_bytes, int64_t* elem_bytes) {
  SGG_CHECK(d && name && offset_bytes, "sgg_ws_lookup: null argument");
  SGG_TRY(check_dims(*d));
  uint8_t* base = reinterpret_cast<uint8_t*>(uintptr_t(1) << 40);
  const Ws w = ws_layout(*d, base);
  struct { const char* n; const void* p; int eb; } tab[] = {
      {"g.X", w.g.X, 2}, {"g.Cf", w.g.Cf, 4}, {"g.CH", w.g.CH, 2}, {"g.EA", w.g.EA, 4}, {"g.Q", w.g.Q, 4},
      {"g.P", w.g.P, 4}, {"g.QB", w.g.QB, 2}, {"g.XB", w.g.XB, 4}, {"g.EB", w.g.EB, 2}, {"g.CB", w.g.CB, 4},
      {"g.PB", w.g.PB, 4},
      {"d.X", w.d.X, 2}, {"d.Cf", w.d.Cf, 4}, {"d.CH", w.d.CH, 2}, {"d.EA", w.d.EA, 4}, {"d.ED", w.d.ED, 4},
      {"d.Q", w.d.Q, 4}, {"d.P", w.d.P, 4}, {"d.QB", w.d.QB, 2}, {"d.XB", w.d.XB, 4}, {"d.EB", w.d.EB, 2},
      {"d.CB", w.d.CB, 4}, {"d.PB", w.d.PB, 4}, {"d.Y", w.d.Y, 4},
      {"FAKE", w.FAKE, 2}, {"DFAKE", w.DFAKE, 4}, {"DFAKEH", w.DFAKEH, 2}, {"VHL", w.VHL, 2}, {"HB", w.HB, 4},
      {"UF", w.UF, 4}, {"UBH", w.UBH, 2}, {"UDB", w.UDB, 2}, {"slopes", w.slopes, 4}, {"coef", w.coef, 4}};
  for (auto& e : tab)
    if (strcmp(e.n, name) == 0) {
      *offset_bytes = (int64_t)((const uint8_t*)e.p - base);
      if (elem_bytes) *elem_bytes = e.eb;
      return 0;
    }
  set_error("sgg_ws_lookup: unknown buffer '%s'", name);
  return -1;
}
