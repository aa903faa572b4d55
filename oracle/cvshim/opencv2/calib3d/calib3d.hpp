// header shim -- TEST INFRASTRUCTURE ONLY: the reference headers include calib3d; nothing of it is used by the
// translation units compiled here
#pragma once
#include <opencv2/core/core.hpp>
