// TEST INFRASTRUCTURE: src/audio.h holds a `json packet` member; src/audio.cpp (the only user) is not compiled.
#pragma once
namespace nlohmann { struct json {}; }
