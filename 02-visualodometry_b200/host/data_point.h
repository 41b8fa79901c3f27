// data_point.h — the measurement / landmark records of the reference (src/data_point.h:6-31) on
// dependency-free field types. Same member names, same constructor argument order.
#pragma once
#include "vo_math.h"

struct Data_Point {
  int id_meas = 0;
  int id_real = 0;
  vo::Point2f coordinates;
  vo::Descriptor descriptor;

  Data_Point() = default;
  Data_Point(int meas_id, int real_id, vo::Point2f coord, const vo::Descriptor& desc)
      : id_meas(meas_id), id_real(real_id), coordinates(coord), descriptor(desc) {}
};

struct World_Point {
  vo::Point3f coordinates;
  vo::Descriptor descriptor;
  int id_real = 0;
  int id_meas = -1;

  World_Point(vo::Point3f coord, const vo::Descriptor& desc, int real_id)
      : coordinates(coord), descriptor(desc), id_real(real_id), id_meas(-1) {}
  World_Point(vo::Point3f coord, const vo::Descriptor& desc, int meas_id, int real_id)
      : coordinates(coord), descriptor(desc), id_real(real_id), id_meas(meas_id) {}
};

using DataPointVector = std::vector<Data_Point>;
using WorldPointVector = std::vector<World_Point>;
