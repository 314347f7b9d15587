// C-ABI plumbing: error strings, device probe, GEMM dispatcher.
#include <stdarg.h>
#include <string.h>

#include "gemm_epilogue.cuh"

static thread_local char g_err[512] = "";

void b200rec_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* b200rec_last_error(void) { return g_err; }
int b200rec_version(void) { return 100; }

int b200rec_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

int gemm_simt_launch(const b200rec_gemm_args* a, const EpiParams& ep, cudaStream_t st);
int gemm_tc_launch(const b200rec_gemm_args* a, const EpiParams& ep, cudaStream_t st);

int gemm_tc_launch_grouped(const b200rec_gemm_args* a, int n, const EpiParams& ep, cudaStream_t st);

static bool nce_args_ok(const b200rec_gemm_args& x) {
  return x.nce_mref && x.nce_thr && x.nce_stats && x.nce_logit_scale && x.c_dtype == B200REC_BF16 &&
         ((uintptr_t)x.nce_stats & 15) == 0 && x.in_dtype == B200REC_BF16;
}

static bool gemm_groupable(const b200rec_gemm_args* a, int n) {
  const b200rec_gemm_args& f = a[0];
  if (f.in_dtype != B200REC_BF16 || (f.epilogue != B200REC_EPI_STORE && f.epilogue != B200REC_EPI_ACCUM &&
                                      f.epilogue != B200REC_EPI_NCE_EXP && f.epilogue != B200REC_EPI_GT_BITS)) return false;
  if (f.epilogue == B200REC_EPI_GT_BITS)       // bit-pack groups: plain threshold only (no rank-1 addend, no row flag)
    for (int g = 0; g < n; ++g)
      if (a[g].gt_row != nullptr || a[g].gt_col != nullptr || a[g].row_scale != nullptr) return false;
  if (f.bias || f.resid || f.n_split != 0 || f.C2 != nullptr) return false;
  for (int g = 0; g < n; ++g) {
    const b200rec_gemm_args& x = a[g];
    if (x.M != f.M || x.N != f.N || x.K != f.K || x.a_major != f.a_major || x.b_major != f.b_major ||
        x.in_dtype != f.in_dtype || x.c_dtype != f.c_dtype || x.epilogue != f.epilogue || x.ldc != f.ldc ||
        x.alpha != f.alpha || x.alpha_dev != f.alpha_dev || x.bias || x.resid || x.n_split != 0 || x.C2 != nullptr ||
        x.nce_logit_scale != f.nce_logit_scale || (x.row_scale == nullptr) != (f.row_scale == nullptr))
      return false;
    if (x.epilogue == B200REC_EPI_NCE_EXP && !nce_args_ok(x)) return false;
    uintptr_t al = x.c_dtype == B200REC_F32 ? 16 : 8;
    if (((uintptr_t)x.C % al) != 0 || ((uintptr_t)x.A & 15) != 0 || ((uintptr_t)x.B & 15) != 0 || x.lda % 8 != 0 ||
        x.ldb % 8 != 0 || x.C == nullptr)
      return false;
  }
  return true;
}

int b200rec_gemm_grouped(const b200rec_gemm_args* a, int n_groups, void* stream) {
  B200_CHECK_ARG(a != nullptr && n_groups >= 0, "gemm_grouped: bad args");
  int done = 0;
  while (done < n_groups) {
    const int n = n_groups - done < 16 ? n_groups - done : 16;
    const b200rec_gemm_args* f = a + done;
    if (n >= 2 && f->M > 0 && f->N > 0 && f->K > 0 && gemm_groupable(f, n)) {
      EpiParams ep;
      memset(&ep, 0, sizeof(ep));
      ep.C = f->C; ep.ldc = f->ldc; ep.c_dtype = f->c_dtype;
      ep.mode = f->epilogue; ep.alpha = f->alpha; ep.alpha_dev = f->alpha_dev;
      ep.M = f->M; ep.N = f->N;
      ep.vec_ok = (f->ldc % 4) == 0;
      ep.fold_id_stride = 1;
      ep.nce_logit_scale = f->nce_logit_scale;
      ep.nce_parts = b200rec_gemm_nce_parts(f->N);
      if (f->epilogue == B200REC_EPI_ACCUM) B200_CHECK_ARG(f->c_dtype == B200REC_F32, "gemm: ACCUM needs fp32 C");
      int rc = gemm_tc_launch_grouped(f, n, ep, (cudaStream_t)stream);
      if (rc) return rc;
    } else {
      // heterogeneous / fp32-verification / fused-epilogue problems: same results, one launch each
      for (int g = 0; g < n; ++g) {
        int rc = b200rec_gemm(f + g, stream);
        if (rc) return rc;
      }
    }
    done += n;
  }
  return 0;
}

int b200rec_gemm(const b200rec_gemm_args* a, void* stream) {
  B200_CHECK_ARG(a != nullptr, "gemm: null args");
  B200_CHECK_ARG(a->M >= 0 && a->N >= 0 && a->K > 0, "gemm: bad shape %d %d %d", a->M, a->N, a->K);
  if (a->M == 0 || a->N == 0) return 0;
  EpiParams ep;
  ep.C = a->C; ep.ldc = a->ldc; ep.c_dtype = a->c_dtype;
  ep.C2 = a->C2; ep.ldc2 = a->ldc2; ep.c2_dtype = a->c2_dtype;
  ep.mode = a->epilogue; ep.alpha = a->alpha; ep.alpha_dev = a->alpha_dev;
  ep.bias = a->bias; ep.resid = a->resid; ep.ldr = a->ldr;
  ep.n_split = a->n_split; ep.c_split_stride = a->c_split_stride; ep.c2_split_stride = a->c2_split_stride;
  ep.M = a->M; ep.N = a->N;
  {
    // vector (4-element) epilogue access is legal when every touched pointer is 16-byte aligned for
    // fp32 / 8-byte for bf16 and every leading dimension is a multiple of 4 elements
    auto ok = [](const void* ptr, int64_t ld, int dtype) {
      if (!ptr) return true;
      uintptr_t al = dtype == B200REC_F32 ? 16 : 8;
      return ((uintptr_t)ptr % al) == 0 && (ld % 4) == 0;
    };
    ep.vec_ok = ok(a->C, a->ldc, a->c_dtype) && ok(a->C2, a->ldc2, a->c2_dtype) && ok(a->bias, 0, B200REC_F32) &&
                ok(a->resid, a->ldr, B200REC_F32) && (a->n_split % 4 == 0) && a->c_split_stride % 4 == 0 &&
                a->c2_split_stride % 4 == 0 && a->epilogue != B200REC_EPI_GT_BITS;
  }
  ep.row_scale = a->row_scale;
  ep.gt_row = a->gt_row; ep.gt_col = a->gt_col;
  B200_CHECK_ARG((a->gt_row == nullptr) == (a->gt_col == nullptr), "gemm: gt_row and gt_col go together");
  B200_CHECK_ARG(a->gt_row == nullptr || (a->epilogue == B200REC_EPI_GT_BITS && a->in_dtype == B200REC_BF16),
                 "gemm: gt_row / gt_col belong to the bf16 GT_BITS epilogue");
  ep.nce_mref = a->nce_mref; ep.nce_thr = a->nce_thr; ep.nce_stats = a->nce_stats;
  ep.nce_logit_scale = a->nce_logit_scale; ep.nce_parts = b200rec_gemm_nce_parts(a->N);
  if (a->epilogue == B200REC_EPI_NCE_EXP)
    B200_CHECK_ARG(nce_args_ok(*a), "gemm: NCE_EXP needs bf16 operands / output and nce_mref, nce_thr, nce_stats "
                                    "(16-byte aligned), nce_logit_scale");
  ep.fold_hp = a->fold_hp; ep.fold_head_on = a->fold_head_on; ep.fold_head_cat = a->fold_head_cat;
  ep.fold_item_tags = a->fold_item_tags; ep.fold_id_offset = a->fold_id_offset; ep.fold_id_stride = a->fold_id_stride;
  ep.fold_thr = a->fold_thr; ep.fold_cnt = a->fold_cnt; ep.fold_keys = (unsigned long long*)a->fold_keys;
  ep.fold_cap = a->fold_cap; ep.fold_groups = a->fold_groups < 1 ? 1 : a->fold_groups;
  if (a->epilogue == B200REC_EPI_FOLD_HEADS) {
    const int hp = a->fold_hp;
    B200_CHECK_ARG(a->in_dtype == B200REC_BF16, "gemm: FOLD_HEADS runs on the tcgen05 path only");
    B200_CHECK_ARG(hp >= 1 && hp <= 32 && (hp & (hp - 1)) == 0 && a->M % hp == 0, "gemm: bad fold_hp %d", hp);
    B200_CHECK_ARG(a->fold_head_on != nullptr && a->fold_id_stride >= 1, "gemm: FOLD_HEADS args");
    B200_CHECK_ARG(a->fold_groups <= 1 || (a->fold_thr != nullptr && a->M % (hp * a->fold_groups) == 0 &&
                                           hp * a->fold_groups <= 32),
                   "gemm: fold_groups > 1 needs the streamed variant, M %% (hp * groups) == 0 and <= 32 heads");
    if (a->fold_thr != nullptr) {
      B200_CHECK_ARG(a->fold_cnt && a->fold_keys && a->fold_cap > 0 && a->N < (1 << 27),
                     "gemm: streamed FOLD_HEADS needs fold_cnt / fold_keys / fold_cap and N < 2^27 rows per shard");
    } else {
      B200_CHECK_ARG(a->C2 != nullptr, "gemm: FOLD_HEADS args");
      B200_CHECK_ARG(a->ldc % 4 == 0 && a->ldc2 % 4 == 0 && ((uintptr_t)a->C % 16) == 0 && ((uintptr_t)a->C2 % 4) == 0,
                     "gemm: FOLD_HEADS needs ldc %% 4 == 0 and aligned outputs");
    }
  }
  ep.fold_on_bits = a->fold_on_bits;
  if (a->epilogue == B200REC_EPI_FOLD_ITEMS) {
    const int hp = a->fold_hp;
    B200_CHECK_ARG(a->in_dtype == B200REC_BF16, "gemm: FOLD_ITEMS runs on the tcgen05 path only");
    B200_CHECK_ARG((hp == 1 || hp == 2 || hp == 4 || hp == 8 || hp == 12 || hp == 16) && a->N % hp == 0,
                   "gemm: FOLD_ITEMS needs fold_hp in {1, 2, 4, 8, 12, 16} and N %% fold_hp == 0 (got %d)", hp);
    B200_CHECK_ARG(a->fold_on_bits != nullptr && a->fold_id_stride >= 1 && a->a_major == 0 && a->b_major == 0,
                   "gemm: FOLD_ITEMS needs fold_on_bits, fold_id_stride >= 1 and K-major operands");
    if (a->fold_thr != nullptr) {
      B200_CHECK_ARG(a->fold_cnt && a->fold_keys && a->fold_cap > 0 && a->M < (1 << 27),
                     "gemm: streamed FOLD_ITEMS needs fold_cnt / fold_keys / fold_cap and M < 2^27 rows per shard");
    } else {
      B200_CHECK_ARG(a->C2 != nullptr && a->ldc >= a->M && a->ldc2 >= a->M, "gemm: FOLD_ITEMS needs C2 and ldc, ldc2 >= M");
    }
  }
  B200_CHECK_ARG(a->C != nullptr || (a->epilogue == B200REC_EPI_FOLD_ITEMS && a->fold_thr != nullptr), "gemm: null C");
  if (a->epilogue == B200REC_EPI_SILU_DUAL) B200_CHECK_ARG(a->C2 != nullptr, "gemm: SILU_DUAL needs C2");
  if (a->epilogue == B200REC_EPI_RESBLOCK) B200_CHECK_ARG(a->resid != nullptr, "gemm: RESBLOCK needs resid");
  if (a->epilogue == B200REC_EPI_ACCUM) B200_CHECK_ARG(a->c_dtype == B200REC_F32, "gemm: ACCUM needs fp32 C");
  if (a->in_dtype == B200REC_F32) return gemm_simt_launch(a, ep, (cudaStream_t)stream);
  if (a->in_dtype == B200REC_BF16) return gemm_tc_launch(a, ep, (cudaStream_t)stream);
  b200rec_set_error("gemm: bad in_dtype %d", a->in_dtype);
  return 1;
}
