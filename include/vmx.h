/*
 * vmx.h -- C ABI of the B200 exponentiation engine for the Verificatum mix-net hot path.
 *
 * This is the drop-in boundary.  It replaces, at ARRAY granularity, the
 * per-element native seam the reference reaches through verificatum-vcr 3.1.0
 * (configure.ac:35): com.verificatum.vmgj.VMG {powm, spowm, fpowm_precomp/fpowm/fpowm_clear,
 * legendre} behind com.verificatum.arithm.{PGroupElementArray, PRingElementArray,
 * LargeIntegerArray, ModPGroup}.  Each entry point cites the reference call site(s) in
 * /root/reference/src/java/com/verificatum/protocol that reach it (paths relative to that
 * directory).  INTEGRATION.md shows the JNI / Panama binding a VCR maintainer would add.
 *
 * Conventions
 *  - Plain C, no exceptions, every function returns a status code; nothing aborts.
 *  - Arrays are opaque DEVICE-resident handles.  Inputs are never consumed or mutated
 *    (Java arrays are immutable values); every returned handle is owned by the caller and
 *    must be released with vmx_garr_free / vmx_rarr_free (the reference's explicit free()
 *    discipline, e.g. hvzk/PoSBasicTW.java:613-656).  A ctx must outlive its arrays.
 *  - Host byte buffers are borrowed for the duration of the call only.
 *  - Wire format of one group element / ring element: fixed-width big-endian two's
 *    complement, `vmx_ctx_elem_bytes` / `vmx_ctx_ring_bytes` bytes (the payload of the
 *    byte-tree leaf the reference writes), elements back to back.
 *  - Inside the engine: 32-bit limbs, vectorised limb-major (uint4 groups of 4 limbs,
 *    element index fastest), group elements resident in Montgomery form, ring elements as
 *    canonical residues.  Conversion happens only in *_from_bytes / *_to_bytes.
 *  - Thread safety: calls on distinct handles may run concurrently; handles are immutable
 *    after creation and may be read concurrently.
 *  - There is NO CPU fallback: without a CUDA device every call fails with VMX_ECUDA.
 */
#ifndef VMX_H
#define VMX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vmx_ctx vmx_ctx;
typedef struct vmx_garr vmx_garr; /* array of group elements (PGroupElementArray over ModPGroup) */
typedef struct vmx_rarr vmx_rarr; /* array of ring elements  (PRingElementArray over Z_q)        */

enum {
  VMX_OK = 0,
  VMX_EFORMAT = 1, /* not in range / not in group / bad length  -> ArithmFormatException */
  VMX_ESIZE = 2,   /* size mismatch between operands            -> ProtocolError          */
  VMX_ENOMEM = 3,
  VMX_ECUDA = 4,   /* no device / kernel failure (see vmx_last_error)                     */
  VMX_EARG = 5
};

const char* vmx_last_error(void);
int vmx_version(void);

/* ---------------------------------------------------------------- context (ModPGroup) */
/* ModPGroup(p, q, g): p odd prime of 2048 or 3072 bits, q | p-1 the (odd) group order,
 * g a generator of the order-q subgroup; all big-endian unsigned, `nbytes` each.
 * Replaces arithm.ModPGroup construction (elgamal/ProtocolElGamal.java:352-434). */
int vmx_ctx_create_modp(const uint8_t* p_be, const uint8_t* q_be, const uint8_t* g_be, size_t nbytes,
                        int device, vmx_ctx** out);
void vmx_ctx_destroy(vmx_ctx* ctx);
size_t vmx_ctx_elem_bytes(const vmx_ctx* ctx); /* bytelen(p) = bits/8 + 1 */
size_t vmx_ctx_ring_bytes(const vmx_ctx* ctx); /* bytelen(q)              */
int vmx_ctx_sync(vmx_ctx* ctx);                /* wait for all queued work */
void* vmx_ctx_stream(vmx_ctx* ctx);            /* the cudaStream_t all work of this ctx is queued on */
/* Window width of fixed-base tables built from now on (0 = choose from the array size). */
int vmx_ctx_set_fixed_window(vmx_ctx* ctx, int w);

/* ---------------------------------------------------------------- group arrays: I/O */
/* PGroup.toElementArray(size, ByteTreeReader) (hvzk/PoSBasicTW.java:507,787-789;
 * mixnet/ShufflerElGamalSession.java:205): parse n fixed-width elements; every element must
 * satisfy 0 < x < p and, if check_membership != 0, x^q = 1 (Legendre symbol 1 for safe primes).
 * On violation returns VMX_EFORMAT and *out = NULL. */
int vmx_garr_from_bytes(vmx_ctx* ctx, size_t n, const uint8_t* be, int check_membership, vmx_garr** out);
/* PGroup.randomElementArray(size, prg, statDist) for ModPGroup (distr/IndependentGeneratorsRO.java:129):
 * element i = (t_i mod p)^((p-1)/q), t_i = the i-th `width`-byte big-endian integer masked to
 * `bitlen` bits.  The bytes come from the caller's PRG; reduction and cofactor power run here. */
int vmx_garr_from_raw(vmx_ctx* ctx, size_t n, const uint8_t* be, size_t width, unsigned bitlen, vmx_garr** out);
/* The same with the bytes drawn on the device from PRGHeuristic(SHA-256) seeded with `seed`
 * (from stream byte `offset`): IndependentGeneratorsRO.generate (distr/IndependentGeneratorsRO.java:110-130;
 * mixnet/ShufflerElGamalSession.java:384; mixnet/MixNetElGamalVerifyFiatShamirSession.java:556-566). */
int vmx_garr_prg_sha256(vmx_ctx* ctx, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t n, size_t width,
                        unsigned bitlen, vmx_garr** out);
/* PGroupElementArray.toByteTree() payload (hvzk/PoSBasicTW.java:694-699). */
int vmx_garr_to_bytes(const vmx_garr* a, uint8_t* be_out);
/* The same two conversions on the byte-tree form itself: `leaves` are the n children of the node
 * an array serialises to, 0x01 || be32(elem_bytes) || payload each (n * (5 + elem_bytes) bytes), so
 * that the host neither strips nor inserts leaf headers (the arrays of a proof are hashed and
 * written in exactly this form, hvzk/PoSTW.java:118-130).  Headers are validated on the device. */
int vmx_garr_from_leaves(vmx_ctx* ctx, size_t n, const uint8_t* leaves, int check_membership, vmx_garr** out);
int vmx_garr_to_leaves(const vmx_garr* a, uint8_t* leaves_out);
/* PGroup.toElementArray(size, PGroupElement): n copies of one element (hvzk/PoSBasicTW.java:805). */
int vmx_garr_fill(vmx_ctx* ctx, size_t n, const uint8_t* elem_be, vmx_garr** out);
void vmx_garr_free(vmx_garr* a);
size_t vmx_garr_size(const vmx_garr* a);

/* ---------------------------------------------------------------- group arrays: algebra */
/* PGroupElement.exp(PRingElementArray): fixed-base, out[i] = base^{e[i]}
 * (mixnet/ShufflerElGamalSession.java:407; hvzk/PoSBasicTW.java:447,606,608,644,646,1030).
 * The window table for `base` is built on first use and cached in the ctx. */
int vmx_exp_fixed(vmx_ctx* ctx, const uint8_t* base_be, const vmx_rarr* e, vmx_garr** out);
/* PGroupElement.exp(PRingElement) on a single element (hvzk/PoSBasicTW.java:481,668,679,690,1014,1021,
 * 1048,1055,1063): uses the cached table of `base` if one exists, else a windowed ladder. */
int vmx_elem_exp(vmx_ctx* ctx, const uint8_t* base_be, const uint8_t* e_be, uint8_t* out_be);
/* PGroupElement.inv() on a single element (hvzk/PoSBasicTW.java:1013-1014, PoSCBasicTW.java:668-669,
 * elgamal/DistrElGamalSessionBasic.java:697,724): O(1) per proof, binary extended Euclid on the host
 * (the reference inverts a host BigInteger here too).  Arrays are inverted by vmx_inv on the device. */
int vmx_elem_inv(vmx_ctx* ctx, const uint8_t* in_be, uint8_t* out_be);
/* Build (or resize) the table of `base` ahead of time for arrays of about n_hint exponents:
 * the analogue of VMG.fpowm_precomp in the reference's native seam. */
int vmx_fixed_precompute(vmx_ctx* ctx, const uint8_t* base_be, size_t n_hint);
/* PGroupElementArray.exp(PRingElementArray): out[i] = a[i]^{e[i]} (hvzk/PoSBasicTW.java:1032). */
int vmx_exp_var(const vmx_garr* a, const vmx_rarr* e, vmx_garr** out);
/* PGroupElementArray.exp(PRingElement): out[i] = a[i]^{e} (hvzk/PoSBasicTW.java:1028;
 * elgamal/DistrElGamalSession.java:384-385; mixnet/PermutationCommitment.java:357). */
int vmx_exp_scalar(const vmx_garr* a, const uint8_t* e_be, vmx_garr** out);
/* PGroupElementArray.expProd(PRingElementArray): prod_i a[i]^{e[i]} -> one element
 * (hvzk/PoSBasicTW.java:408-409,481,690,1021,1063; hvzk/CCPoSBasicW.java:381,394,499-504;
 * elgamal/DistrElGamalSessionBasic.java:524-526,683-685,707-709).
 * The `k` arrays share the exponents (a product-group array = k component arrays). */
int vmx_expprod(const vmx_garr* const* a, size_t k, const vmx_rarr* e, uint8_t* out_be /* k elements */);
/* PGroup.expProd(PGroupElementArray[] bases, LargeInteger[] integers, bitLength): element-wise
 * out[i] = prod_j bases[j][i]^{ints[j]}, small signed integers
 * (elgamal/DistrElGamalSessionBasic.java:502). */
int vmx_expprod_cols(const vmx_garr* const* bases, size_t t, const int64_t* ints, vmx_garr** out);
/* PGroupElementArray.mul (mixnet/ShufflerElGamalSession.java:273; hvzk/PoSBasicTW.java:448,610). */
int vmx_mul(const vmx_garr* a, const vmx_garr* b, vmx_garr** out);
/* PGroupElementArray.inv (used by div of arrays). */
int vmx_inv(const vmx_garr* a, vmx_garr** out);
/* PGroupElementArray.prod() (hvzk/PoSBasicTW.java:1013). */
int vmx_prod(const vmx_garr* a, uint8_t* out_be);
/* PGroupElementArray.permute(Permutation): out[perm[i]] = a[i]
 * (mixnet/ShufflerElGamalSession.java:278; hvzk/PoSBasicTW.java:451). */
int vmx_permute(const vmx_garr* a, const uint32_t* perm, vmx_garr** out);
/* PGroupElementArray.shiftPush(el): out[0] = el, out[i] = a[i-1] (hvzk/PoSBasicTW.java:1031). */
int vmx_shift_push(const vmx_garr* a, const uint8_t* elem_be, vmx_garr** out);
/* PGroupElementArray.extract(boolean[]) (mixnet/PermutationCommitment.java:462-468). */
int vmx_extract(const vmx_garr* a, const uint8_t* keep, vmx_garr** out);
/* PGroupElementArray.copyOfRange(a, b) (hvzk/PoSBasicTW.java:512). */
int vmx_slice(const vmx_garr* a, size_t begin, size_t end, vmx_garr** out);
/* PGroupElementArray.equals (hvzk/PoSBasicTW.java:1035): *equal = 1/0. */
int vmx_equals(const vmx_garr* a, const vmx_garr* b, int* equal);
/* PGroupElementArray.get(i) (hvzk/PoSBasicTW.java:562,1014). */
int vmx_get(const vmx_garr* a, size_t i, uint8_t* out_be);

/* ---------------------------------------------------------------- exchange between GPUs
 * The N ciphertexts are sharded over the GPUs of one box in contiguous index ranges; a
 * permutation (mixnet/ShufflerElGamalSession.java:278, hvzk/PoSBasicTW.java:451,553) moves elements
 * between shards.  pack_rows writes the selected elements element-major (count rows of
 * vmx_ctx_row_bytes bytes, internal representation: Montgomery form for group arrays) into a
 * caller-provided DEVICE buffer that NCCL all-to-all moves over NVLink; unpack_rows builds a fresh
 * array of n elements from such rows (dst_idx must be a permutation of 0..n-1; NULL = identity).
 * idx / dst_idx are HOST lists.  Work is queued on the ctx stream (vmx_ctx_stream). */
size_t vmx_ctx_row_bytes(const vmx_ctx* ctx);
int vmx_garr_pack_rows(const vmx_garr* a, const uint32_t* idx, size_t count, void* rows_dev);
int vmx_garr_unpack_rows(vmx_ctx* ctx, size_t n, const void* rows_dev, const uint32_t* dst_idx, size_t count, vmx_garr** out);
int vmx_rarr_pack_rows(const vmx_rarr* a, const uint32_t* idx, size_t count, void* rows_dev);
int vmx_rarr_unpack_rows(vmx_ctx* ctx, size_t n, const void* rows_dev, const uint32_t* dst_idx, size_t count, vmx_rarr** out);

/* ---------------------------------------------------------------- ring arrays (Z_q) */
/* PRing/PField.toElementArray(size, ByteTreeReader) (hvzk/PoSBasicTW.java:977,980): each
 * element must satisfy 0 <= x < q, else VMX_EFORMAT. */
int vmx_rarr_from_bytes(vmx_ctx* ctx, size_t n, const uint8_t* be, vmx_rarr** out);
/* Raw integers of `width` bytes each (big-endian, unsigned), reduced mod q:
 * LargeIntegerArray.random(size, bitlen, randomSource) + pField.toElementArray
 * (hvzk/PoSBasicTW.java:472-474) and pRing.randomElementArray (:446,571,621) -- the random
 * bytes stay owned by the caller's RandomSource; `bitlen` masks the top bits (0 = keep all). */
int vmx_rarr_from_raw(vmx_ctx* ctx, size_t n, const uint8_t* be, size_t width, unsigned bitlen, vmx_rarr** out);
/* PRG-derived batching vector: prg.setSeed(seed); LargeIntegerArray.random(n, bitlen, prg)
 * with PRGHeuristic(SHA-256) (hvzk/PoSBasicTW.java:533-538; same in PoSCBasicTW.java:350-355,
 * CCPoSBasicW.java:330-335, elgamal/DistrElGamalSessionBasic.java:513-518). */
int vmx_rarr_prg_sha256(vmx_ctx* ctx, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t n, unsigned bitlen,
                        vmx_rarr** out);
/* pRing.randomElementArray(size, prg, statDist) with a PRGHeuristic(SHA-256) source at stream byte `offset`:
 * element i = (i-th `width`-byte integer masked to `bitlen` bits) mod q, bytes drawn on the device. */
int vmx_rarr_prg_raw_sha256(vmx_ctx* ctx, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t n,
                            size_t width, unsigned bitlen, vmx_rarr** out);
/* Bytes [offset, offset + nbytes) of the PRGHeuristic(SHA-256) stream of `seed`, expanded on the device and
 * copied to the host (Permutation.random, mixnet/ShufflerElGamalSession.java:408-409, PermutationCommitment.java:211). */
int vmx_prg_bytes_sha256(vmx_ctx* ctx, const uint8_t* seed, size_t seedlen, uint64_t offset, size_t nbytes, uint8_t* out);
int vmx_rarr_to_bytes(const vmx_rarr* a, uint8_t* be_out);
int vmx_rarr_from_leaves(vmx_ctx* ctx, size_t n, const uint8_t* leaves, vmx_rarr** out); /* as vmx_garr_from_leaves */
int vmx_rarr_to_leaves(const vmx_rarr* a, uint8_t* leaves_out);
int vmx_rarr_fill(vmx_ctx* ctx, size_t n, const uint8_t* elem_be, vmx_rarr** out);
void vmx_rarr_free(vmx_rarr* a);
size_t vmx_rarr_size(const vmx_rarr* a);
/* maximal bit length over the array (LargeIntegerArray.bitLength analogue; cached). */
int vmx_rarr_bitlen(const vmx_rarr* a, unsigned* bits);

int vmx_radd(const vmx_rarr* a, const vmx_rarr* b, vmx_rarr** out);             /* hvzk/PoSBasicTW.java:643 */
int vmx_rneg(const vmx_rarr* a, vmx_rarr** out);
int vmx_rsub(const vmx_rarr* a, const vmx_rarr* b, vmx_rarr** out);
int vmx_rmul(const vmx_rarr* a, const vmx_rarr* b, vmx_rarr** out);             /* :642,645 */
/* PRingElementArray.mulAdd(scalar, arr): out[i] = a[i]*s + b[i] (:874,877). */
int vmx_rmuladd(const vmx_rarr* a, const uint8_t* s_be, const vmx_rarr* b, vmx_rarr** out);
int vmx_rinner(const vmx_rarr* a, const vmx_rarr* b, uint8_t* out_be);          /* :861,863 */
int vmx_rsum(const vmx_rarr* a, uint8_t* out_be);                               /* :862 */
int vmx_rprod(const vmx_rarr* a, uint8_t* out_be);                              /* :1014 e.prod() */
/* PRingElementArray.prods(): out[i] = a[0]*...*a[i] (:604). */
int vmx_rprods(const vmx_rarr* a, vmx_rarr** out);
/* PRingElementArray.recLin(e): x[0] = b[0], x[i] = x[i-1]*e[i] + b[i]; returns x and d = x[n-1]
 * (:583-598). */
int vmx_rreclin(const vmx_rarr* b, const vmx_rarr* e, vmx_rarr** out, uint8_t* last_be);
int vmx_rpermute(const vmx_rarr* a, const uint32_t* perm, vmx_rarr** out);      /* :553 */
int vmx_rshift_push(const vmx_rarr* a, const uint8_t* elem_be, vmx_rarr** out); /* :637-638 */
int vmx_rslice(const vmx_rarr* a, size_t begin, size_t end, vmx_rarr** out);
int vmx_rget(const vmx_rarr* a, size_t i, uint8_t* out_be);
int vmx_requals(const vmx_rarr* a, const vmx_rarr* b, int* equal);

/* ---------------------------------------------------------------- instrumentation */
/* Number of engine kernels launched on this ctx since creation (bench.py `gpu_launches`). */
uint64_t vmx_ctx_launch_count(const vmx_ctx* ctx);
/* Modular multiplications executed by the engine's kernels since creation (host-side
 * accounting of the work each launch performs; used for the roofline figure). */
uint64_t vmx_ctx_modmul_count(const vmx_ctx* ctx);
/* Raw batched Montgomery multiplication benchmark kernel: out[i] = a[i]*b[i]^iters (same code
 * path as every exponentiation); returns elapsed device ms through *ms. */
int vmx_bench_modmul(vmx_ctx* ctx, size_t n, int iters, float* ms);

/* Self test: a[i]*b[i] through the thread-per-element and the warp-cooperative multiplier must
 * agree word for word (*equal = 1). */
int vmx_selftest_coop(const vmx_garr* a, const vmx_garr* b, int* equal);
/* Test hook: out[i] = a[i]*b[i] computed by the warp-cooperative multiplier. */
int vmx_debug_coop_mul(const vmx_garr* a, const vmx_garr* b, vmx_garr** out);

#ifdef __cplusplus
}
#endif
#endif /* VMX_H */
