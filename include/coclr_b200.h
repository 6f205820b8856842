/*
 * coclr_b200 -- C ABI of the B200 (sm_100a) kernels behind the CoCLR data-parallel hot path.
 *
 * The reference (TengdaHan/CoCLR) has no FFI layer: its hot path is Python calling nn.Conv3d /
 * nn.BatchNorm3d / nn.MaxPool3d / torch.einsum.  Each entry point below
 * replaces the vendor-library call(s) named in its comment (reference file:line), takes raw device
 * pointers + sizes + a cudaStream_t, never allocates, never synchronises, and returns 0 on success
 * or a negative error code (COCLR_E_*).  All activations are channels-last (N,D,H,W,C) fp32 rows.
 *
 * The Python host side (coclr_b200/lib.py) binds these with ctypes; INTEGRATION.md shows the stub.
 */
#ifndef COCLR_B200_H_
#define COCLR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COCLR_OK 0
#define COCLR_E_ARG (-1)    /* bad argument (shape / alignment / null) */
#define COCLR_E_LAUNCH (-2) /* CUDA launch error (cudaGetLastError != success) */
#define COCLR_E_DEVICE (-3) /* not an sm_100 device */

typedef void* coclr_stream_t; /* cudaStream_t */

/* A channels-last fp32 activation as an implicit-GEMM operand source.  The value used is
 *   relu?( scale[c] * x + shift[c] )   (BatchNorm-apply + ReLU folded into the operand load),
 * or x itself when scale == NULL. */
typedef struct {
  const float* ptr;   /* [B, T, H, W, ld] */
  int ld;             /* channels per pixel in memory */
  int coff;           /* first channel used */
  int C;              /* channels used (multiple of 4) */
  int T, H, W;
  const float* scale; /* [C] or NULL */
  const float* shift; /* [C] or NULL */
  int relu;
} coclr_src_t;

typedef struct {
  int kt, kh, kw;
  int st, sh, sw; /* each 1 or 2 */
  int pt, ph, pw;
  int transposed; /* 0: src = dst*s + k - p (forward conv / wgrad gather); 1: src = (dst + p - k)/s (dgrad) */
} coclr_geom_t;

/* ---- implicit-GEMM convolution, forward and data-gradient -----------------------------------
 * replaces nn.Conv3d forward (backbone/s3dg.py:11-13,39-42; model/pretrain.py:52,54) and the cuDNN
 * dgrad reached through loss.backward() (main_nce.py:330).  Y[m, n] (+)= sum_k A[m, k] * Wp[n, k],
 * m over the B*Td*Hd*Wd destination pixels, k = tap * C + channel.  Optionally accumulates the
 * per-channel sum / sum of squares of Y (train-mode BatchNorm3d statistics, s3dg.py:16,46-47). */
typedef struct {
  coclr_src_t src;
  coclr_geom_t g;
  int B, Td, Hd, Wd; /* destination pixel grid */
  int Kreal;         /* taps * src.C */
  const void* wpk;   /* packed weights from coclr_pack_weights */
  const float* wunscale; /* [n_tiles*BN] per-column 1/scale, or NULL */
  int N, BN, n_tiles;    /* real output channels, tile width (multiple of 32, <= 256), tiles */
  float* dst;
  int dst_ld, dst_coff;
  int accumulate;  /* dst += result */
  double* stats;   /* [2*N]: sum, then sum of squares; NULL to skip */
  int npass;       /* 1 = single 16-bit pass, 3 = hi/lo split (fp32-equivalent) */
  int bf16;        /* 0 = fp16 operands, 1 = bf16 operands */
} coclr_conv_t;
int coclr_conv_igemm(const coclr_conv_t* p, int num_sms, coclr_stream_t stream);
size_t coclr_conv_packed_bytes(int N, int Kreal, int* BN_out, int* n_tiles_out);

/* ---- weight-gradient -----------------------------------------------------------------------
 * replaces cuDNN wgrad.  dW[n, c, tap] += sum_m dY[m, n] * A[m, (tap, c)] into the PyTorch weight
 * layout [Cout, Cin_real, kt, kh, kw] (fp32 atomics; zero dW first). */
typedef struct {
  coclr_src_t src;   /* conv input (forward operand, affine+relu folded) */
  coclr_geom_t g;    /* forward geometry (transposed = 0) */
  coclr_src_t dy;    /* output gradient [B,Td,Hd,Wd,ld], C = Cout rounded up to 4; scale = NULL */
  int B, Td, Hd, Wd;
  int Cout, Cin_real; /* real sizes of dW (src.C may be padded, e.g. 4 for the RGB stem) */
  float* dw;
  int npass, bf16;
  int splits;        /* pixel-range splits (>= 1) */
} coclr_wgrad_t;
int coclr_conv_wgrad(const coclr_wgrad_t* p, coclr_stream_t stream);

/* ---- weight packing -------------------------------------------------------------------------
 * PyTorch conv weight [Cout, Cin, kt, kh, kw] -> swizzled 16-bit hi/lo tile images.
 * mode 0 (forward): rows n = cout, k = tap*Cpad + cin.   mode 1 (dgrad): rows n = cin, k = tap*Cout_pad + cout.
 * Each row is scaled by a power of two (max |w| -> [0.5,1)) before the split; 1/scale goes to unscale. */
typedef struct {
  const float* w;
  int Cout, Cin, taps;
  int Cpad;      /* padded channel count of the K index (>= Cin for mode 0, >= Cout for mode 1) */
  int mode;
  int bf16;
  void* wpk;
  float* unscale; /* [n_tiles*BN] */
} coclr_pack_t;
int coclr_pack_weights(const coclr_pack_t* p, coclr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COCLR_B200_H_ */
