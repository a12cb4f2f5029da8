// spk_shape.cu -- one shape-specialised instantiation of the fused kernels; compiled once per line of spk_shapes.def
// with -DSPK_SHAPE_NX=.. -DSPK_SHAPE_NU=.. -DSPK_SHAPE_D=.. -DSPK_SHAPE_TB=.. -DSPK_SHAPE_TF=..
#define SPK_SHAPE_TU 1
#include "spk_shapes.h"

#define SPK_CAT_(a, b, c, d) a##b##_##c##_##d
#define SPK_CAT(a, b, c, d) SPK_CAT_(a, b, c, d)

extern "C" const SpkShapeEntry *SPK_CAT(spk_shape_entry_, SPK_SHAPE_NX, SPK_SHAPE_NU, SPK_SHAPE_D)() {
  constexpr int NX = SPK_SHAPE_NX, NU = SPK_SHAPE_NU, DD = SPK_SHAPE_D;
  static const SpkShapeEntry e = {
      NX, NU, DD, SPK_SHAPE_TB, SPK_SHAPE_TF,
      spk_k_bwd<NX, NU, DD, SPK_SHAPE_TB>, spk_k_fwd_s1<NX, NU, DD, SPK_SHAPE_TF>, spk_k_dual<NX, NU, DD>, spk_k_dual2<NX, NU, DD>,
#define SPK_D3(LV, UC) {spk_k_dual3<NX, NU, DD, LV, UC, false>, spk_k_dual3<NX, NU, DD, LV, UC, true>}
      {{SPK_D3(false, false), SPK_D3(false, true)}, {SPK_D3(true, false), SPK_D3(true, true)}},
#undef SPK_D3
      {{spk_k_dual3<NX, NU, DD, false, false, false, true>, spk_k_dual3<NX, NU, DD, false, false, true, true>},
       {spk_k_dual3<NX, NU, DD, false, true, false, true>, spk_k_dual3<NX, NU, DD, false, true, true, true>}},
      {spk_k_bwd_mm<NX, NU, DD, false>, spk_k_bwd_mm<NX, NU, DD, true>}, spk_k_fwd_mm<NX, NU, DD>, spk_k_top_mm<NX, NU, DD>,
      spk_k_bwd_top<NX, NU, DD, SPK_SHAPE_TB>, spk_k_fwd_top<NX, NU, DD, SPK_SHAPE_TF>};
  return &e;
}
