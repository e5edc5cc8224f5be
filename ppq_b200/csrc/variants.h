// variants.h -- run-time kernel variant selection (profiling / A-B benchmarking knob, see ppq_b200_set_variant).
#pragma once
namespace ppqb {
enum { kVarLinearT = 0, kVarHistogram = 1, kVarMinMax = 2, kVarCount = 8 };
int variant_of(int key);
}
