// stub: the extractor includes this header but uses nothing from it (orbextractor.hpp:63)
#pragma once
