// opencv2/core.hpp of the type shim (see vo_cv_shim.h)
#pragma once
#include "../vo_cv_shim.h"
