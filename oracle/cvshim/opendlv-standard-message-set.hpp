// stub: the extractor includes this header but uses nothing from it (orbextractor.hpp:64)
#pragma once
