// opencv2/core/eigen.hpp of the type shim (see vo_cv_shim.h)
#pragma once
#include "../../vo_cv_shim.h"
