// vo_records.h — the measurement / landmark records that cross the host interface, on dependency-free field
// types (vo_math.h).  They mirror the reference's records (src/data_point.h:6-31) member for member - same names,
// same constructor argument order - because the drivers (exec/icp_test.cpp, exec/vo.cpp) and match_points<> access
// the members directly; what is added here is what the GPU shims need on top: flat gathers into the row-major
// arrays the C ABI takes (include/vo_b200.h), so that the per-record heap descriptors of the reference are walked
// exactly once per call.
#pragma once
#include <cstdint>
#include <vector>

#include "vo_math.h"

// One image measurement of one frame: `point <id_meas> <id_real> <u> <v> <10 descriptor floats>` in data/meas-*.dat
// (src/my_utilities.cpp:35-112).
struct Data_Point {
  int id_meas = 0;            // index inside the frame
  int id_real = 0;            // ground-truth landmark id (evaluation only)
  vo::Point2f coordinates;    // pixel (u, v)
  vo::Descriptor descriptor;  // 10 floats in the bundled dataset; any dimension 1..16 is accepted by vo_match

  Data_Point() = default;
  Data_Point(int meas_id, int real_id, vo::Point2f coord, const vo::Descriptor& desc)
      : id_meas(meas_id), id_real(real_id), coordinates(coord), descriptor(desc) {}
};

// One landmark of the map: triangulated position + the descriptor and ids of the FIRST view it was seen in
// (src/cam.cpp:122-139).
struct World_Point {
  vo::Point3f coordinates;
  vo::Descriptor descriptor;
  int id_real = 0;
  int id_meas = -1;  // -1: not tied to a measurement (the three-argument constructor of the reference)

  World_Point(vo::Point3f coord, const vo::Descriptor& desc, int real_id)
      : coordinates(coord), descriptor(desc), id_real(real_id), id_meas(-1) {}
  World_Point(vo::Point3f coord, const vo::Descriptor& desc, int meas_id, int real_id)
      : coordinates(coord), descriptor(desc), id_real(real_id), id_meas(meas_id) {}
};

using DataPointVector = std::vector<Data_Point>;
using WorldPointVector = std::vector<World_Point>;

namespace vo {

// descriptor dimension of a record set (0 when empty); every record must agree
template <class Record>
inline int descriptor_dim(const std::vector<Record>& v) {
  return v.empty() ? 0 : static_cast<int>(v.front().descriptor.size());
}

// row-major float[N][D] copy of the descriptors, the layout vo_match takes
template <class Record>
inline void gather_descriptors(const std::vector<Record>& v, std::vector<float>& out) {
  const int d = descriptor_dim(v);
  out.resize(v.size() * static_cast<size_t>(d));
  float* dst = out.data();
  for (const Record& r : v)
    for (int k = 0; k < d; ++k) *dst++ = r.descriptor[k];
}

// the id_real column, for the statistics line match_points prints (src/my_utilities.h:116-119)
template <class Record>
inline void gather_real_ids(const std::vector<Record>& v, std::vector<int32_t>& out) {
  out.resize(v.size());
  for (size_t i = 0; i < v.size(); ++i) out[i] = static_cast<int32_t>(v[i].id_real);
}

}  // namespace vo
