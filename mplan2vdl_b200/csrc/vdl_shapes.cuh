// Static shapes of the fused scan: compile-time answers to the structural questions of a descriptor
// (see "SHAPES" in vdl_fused.cu).  A shape only fixes STRUCTURE -- counts, column widths, which of
// shr / a / b are trivial, what fits 32 bits; every constant stays a run-time descriptor field, so a shape
// serves any plan of that structure (other columns, other bounds, other scale factors).  The host checks all
// of a shape's assumptions against the prepared descriptor (shape_matches) before launching it and otherwise
// uses GenericShape; results are identical.  tools/print_shape.py prints the struct for a plan (needs a GPU:
// narrowing depends on the column statistics).
#pragma once

#define FF_COLN (FF_SHR0 | FF_B1 | FF_A0 | FF_NARROW)      /* a plain 8-byte column whose values fit int32 */

// select(3 ranges: int32 col, 2 x narrow 8-byte col) -> SUM(col * col), COUNT          [TPC-H Q6, Vlite.hs:721-730]
struct ShapeSel3Sum2 {
  static constexpr bool kStatic = true;
  static constexpr const char *kName = "sel3_sum2";
  static constexpr int NPREDS = 3, NKEYS = 0, NACC = 2, KEY32 = 0;
  static constexpr int PRED_MODE[VDL_MAX_PREDS] = {1, 2, 2}, PRED_SHR0[VDL_MAX_PREDS] = {1, 1, 1};
  static constexpr int KEY_FLAGS[VDL_MAX_KEYS] = {}, KEY_SHL0[VDL_MAX_KEYS] = {};
  static constexpr int ACC_OP[K_MAX_ACC] = {0, 0}, ACC_CHAIN[K_MAX_ACC] = {0, 0}, ACC_NFAC[K_MAX_ACC] = {2, 0};
  static constexpr int FAC[K_MAX_ACC][VDL_MAX_FACTORS] = {{FF_COLN, FF_COLN}, {}};
};

// select(1 range on an int32 col) -> group by a 2-part narrow key -> SUM c, SUM c, SUM c*(k-c), chain*(k+c), SUM c,
// COUNT, first row                                                                         [TPC-H Q1, Vlite.hs:1048-1098]
struct ShapeSel1Key2Sum5 {
  static constexpr bool kStatic = true;
  static constexpr const char *kName = "sel1_key2_sum5";
  static constexpr int NPREDS = 1, NKEYS = 2, NACC = 7, KEY32 = 1;
  static constexpr int PRED_MODE[VDL_MAX_PREDS] = {1}, PRED_SHR0[VDL_MAX_PREDS] = {1};
  static constexpr int KEY_FLAGS[VDL_MAX_KEYS] = {FF_NARROW, FF_NARROW}, KEY_SHL0[VDL_MAX_KEYS] = {1, 1};
  static constexpr int ACC_OP[K_MAX_ACC] = {0, 0, 0, 0, 0, 0, 1};
  static constexpr int ACC_CHAIN[K_MAX_ACC] = {0, 0, 1, 1, 0, 0, 0};
  static constexpr int ACC_NFAC[K_MAX_ACC] = {1, 1, 1, 1, 1, 0, 1};
  static constexpr int FAC[K_MAX_ACC][VDL_MAX_FACTORS] = {
      {FF_COLN}, {FF_COLN}, {FF_SHR0 | FF_BM1 | FF_NARROW}, {FF_SHR0 | FF_B1 | FF_NARROW}, {FF_COLN}, {},
      {FF_ROWID | FF_SHR0 | FF_B1 | FF_A0}};
};
