// data_point.h — drop-in for the reference's src/data_point.h:6-31: the same records, member for member (the drivers
// and match_points<> access the fields directly), on the type shims.
#pragma once
#include <Eigen/Core>
#include <opencv2/core.hpp>
#include <vector>

struct Data_Point {
  int id_meas;
  int id_real;
  cv::Point2f coordinates;
  Eigen::VectorXf descriptor;
  Data_Point(int meas_id, int real_id, cv::Point2f coord, const Eigen::VectorXf& desc)
      : id_meas(meas_id), id_real(real_id), coordinates(coord), descriptor(desc) {}
  Data_Point() : id_meas(0), id_real(0), coordinates(0.0f, 0.0f), descriptor(Eigen::VectorXf()) {}
};

struct World_Point {
  cv::Point3f coordinates;
  Eigen::VectorXf descriptor;
  int id_real;
  int id_meas;
  World_Point(cv::Point3f coord, const Eigen::VectorXf& desc, int real_id)
      : coordinates(coord), descriptor(desc), id_real(real_id), id_meas(-1) {}
  World_Point(cv::Point3f coord, const Eigen::VectorXf& desc, int meas_id, int real_id)
      : coordinates(coord), descriptor(desc), id_real(real_id), id_meas(meas_id) {}
};

typedef std::vector<Data_Point> DataPointVector;
typedef std::vector<World_Point> WorldPointVector;
