/* vmnv.h -- C ABI of the native universal verifier (verificatum-vmn_b200/libvmnv.so, csrc/vmnv_native.cpp).
 *
 * What `vmnv` does with the proof directory of a mix-net execution
 * (mixnet/MixNetElGamalVerifyFiatShamirSession.java:1318-1668: proofs of type "mixing", "shuffling" or "decryption",
 * with or without pre-computation, over a ModPGroup or an ECqPGroup, ciphertexts of any width): every file is walked
 * here, every group / ring operation is one call into the engine's C ABI (include/vmx.h), Fiat-Shamir hashing is
 * SHA-256 on a worker thread beside the GPU.  A JVM would bind these
 * three functions the way INTEGRATION.md binds vmx.h (JNI or FFM); the parameters are what the protocol info file
 * holds (elgamal/ProtocolElGamalGen.java:81-213).
 */
#ifndef VMNV_H
#define VMNV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vmxv_params {
  int kind;            /* 0: ModPGroup (p, q, g); 1: ECqPGroup y^2 = x^3 + a x + b over F_p, generator (g, gy) of order q */
  const uint8_t* p_be; /* big-endian, nbytes each */
  const uint8_t* q_be;
  const uint8_t* g_be; /* the generator (its x coordinate on a curve) */
  const uint8_t* a_be; /* curve only (else NULL) */
  const uint8_t* b_be;
  const uint8_t* gy_be;
  size_t nbytes;
  int device;                 /* CUDA device of the engine context */
  int k, threshold;           /* parties, threshold */
  int vbitlenro, ebitlenro, rbitlen;
  const char* version;        /* VCR.version() as written to the `version` file */
  const char* sid;            /* session identifier of the info file; the auxsid is read from the proof */
  const char* pgroup_string;  /* the pgroup string of the info file (enters the global prefix) */
  const char* expected_auxsid; /* NULL or "": accept the one in the proof (vmnv -auxsid) */
  int expected_width;          /* > 0: the width the proof must have -- vmnv's `-width w`, or, for its default behaviour,
                                  the `width` of the protocol info file (determineWidth :404-440: without the option the
                                  proof's width must equal the info file's); <= 0: accept the one in the proof */
  const char* expected_type;   /* NULL or "": accept the type in the proof; else "mixing" | "shuffling" | "decryption"
                                  (vmnv -mix / -shuffle / -decrypt; MixNetElGamalVerifyFiatShamirSession.java:329-358) */
  int nodec, noposc, noccpos;  /* non-zero: do not verify the decryption / the proofs of shuffles of commitments / the
                                  (commitment-consistent) proofs of shuffles (vmnv -nodec, -noposc, -noccpos;
                                  mixnet/SessionParams.java) */
} vmxv_params;

typedef struct vmxv_file {
  const char* name;    /* path relative to the proof directory, '/' separated: "proofs/PoSReply01.bt" */
  const uint8_t* data; /* borrowed for the call (page-locked memory makes the imports DMA transfers) */
  size_t size;
} vmxv_file;

typedef struct vmxv_report {
  int accepted;      /* the verdict of vmnv */
  int fail_stop;     /* 1: a condition under which the reference stops with an error (`error` says which) */
  int type;          /* 0 "mixing", 1 "shuffling", 2 "decryption" (the `type` file) */
  int n_shuffles;    /* index of the last party whose shuffle was looked at */
  int shuffles[64];  /* party l - 1: 1 its shuffle is valid, -1 invalid (its output is replaced by its input), 0 the
                        party took no part / shuffles are not verified.  After a pre-computation the verdict is that of
                        the proof of a shuffle of commitments AND of the commitment-consistent proof of a shuffle */
  int poscs[64];     /* pre-computation only: 1 / -1 / 0 for the proof of a shuffle of commitments alone (an invalid
                        one replaces the permutation commitment by the generators) */
  int valid_proofs;
  int enough_valid_proofs; /* valid_proofs >= threshold (1 also when no shuffle is verified) */
  int decryption;    /* the combined proof of the decryption factors: 1 valid, 0 invalid, -1 not verified */
  int plaintexts;    /* Plaintexts.bt equals the decrypted output: 1 / 0 / -1 */
  uint64_t hashed_bytes, launches;
  char error[400];
  char test_vectors[16384]; /* the scalar test vectors of `vmnv -t` (mixnet/MixNetElGamalVerifyFiatShamirTool.java:82-224:
                               par.*, der.rho, PoS.s / PoS.v, PoSC.*, CCPoS.*, Dec.s / Dec.v) in the order the reference
                               prints them, one per line: name '@' party (0: none) '=' value (seeds and rho in
                               hexadecimal, challenges and parameters in decimal); cut off when the buffer is full */
} vmxv_report;

/* Bind the engine (path of libvmx.so, or of the host-emulation build in the CPU tests).  0 on success. */
int vmxv_bind(const char* libvmx_path);
/* Verify the proof directory given as in-memory files.  Returns 0 when `report` is filled (also for a rejected or
 * fail-stopped proof), negative on an engine or usage error (report->error). */
int vmxv_verify(const vmxv_params* params, const vmxv_file* files, size_t nfiles, vmxv_report* report);

#ifdef __cplusplus
}
#endif
#endif
