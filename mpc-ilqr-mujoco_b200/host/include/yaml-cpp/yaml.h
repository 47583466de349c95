// Placeholder so that `#include <yaml-cpp/yaml.h>` in code written against the reference's headers resolves;
// the YAML-subset reader lives in src/config.cpp (yaml-cpp is not installed in this image).
#pragma once
