// Frozen record layout of the reference's star_t / binary_t (reference include/cmc/cmc.h:409-601, :231-390).
// The drop-in keeps the reference's AoS arrays byte for byte (restart files fwrite them raw, reference
// src/cmc/cmc_io.c:3862-3867), so the kernels address fields by these offsets.  tests/test_layout.py checks every
// constant against offsetof() in the compiled reference.
#pragma once

#define CMCB_SIZEOF_STAR 424
#define CMCB_SIZEOF_BINARY 464
#define CMCB_STAR_WORDS (CMCB_SIZEOF_STAR / 8)     /* 53 */
#define CMCB_BINARY_WORDS (CMCB_SIZEOF_BINARY / 8) /* 58 */

#define CMCB_OFF_R 0
#define CMCB_OFF_VR 8
#define CMCB_OFF_VT 16
#define CMCB_OFF_M 24
#define CMCB_OFF_E 32
#define CMCB_OFF_J 40
#define CMCB_OFF_EI 48
#define CMCB_OFF_EINT 56
#define CMCB_OFF_RNEW 64
#define CMCB_OFF_VRNEW 72
#define CMCB_OFF_VTNEW 80
#define CMCB_OFF_ROLD 88
#define CMCB_OFF_X 96
#define CMCB_OFF_Y 104
#define CMCB_OFF_R_PERI 112
#define CMCB_OFF_R_APO 120
#define CMCB_OFF_PHI 128
#define CMCB_OFF_INTERACTED 136
#define CMCB_OFF_BININD 152
#define CMCB_OFF_ID 160
#define CMCB_OFF_UOLDROLD 176
#define CMCB_OFF_UOLDRNEW 184
#define CMCB_OFF_VTOLD 192
#define CMCB_OFF_VROLD 200
#define CMCB_OFF_THREEBB_INTERACTED 144
#define CMCB_OFF_RAD 168
#define CMCB_OFF_SE_K 224 /* int */
#define CMCB_BIN_OFF_A 64
