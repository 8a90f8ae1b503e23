// TEST INFRASTRUCTURE ONLY -- see gdal_shim.hpp
#pragma once
#include "gdal_shim.hpp"
