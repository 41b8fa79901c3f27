// data_point.h — kept so that sources written against the reference keep their include line; the records live
// in vo_records.h.
#pragma once
#include "vo_records.h"
