// Umbrella header (reference: loam/loam.h:7-11).
#pragma once
#include "loam/common.h"
#include "loam/features.h"
#include "loam/geometry.h"
#include "loam/registration.h"
