// TEST INFRASTRUCTURE ONLY (oracle build). Link stub P5 of SURVEY.md section 8c:
// the reference declares CPUAdapter (gpu_adaptability.hpp:538-573) and takes its
// typeinfo at gpu_adaptability.cpp:611 but never defines its virtuals, so the
// reference cannot link. These bodies are never reached by the CPU time-stepping path.
#include "weather_sim/gpu_adaptability.hpp"

namespace weather_sim {
bool CPUAdapter::initialize(int) { return true; }
double CPUAdapter::executeShallowWaterStep(const WeatherGrid&, WeatherGrid&, scalar_t) { return 0.0; }
double CPUAdapter::executeBarotropicStep(const WeatherGrid&, WeatherGrid&, scalar_t) { return 0.0; }
double CPUAdapter::executePrimitiveEquationsStep(const WeatherGrid&, WeatherGrid&, scalar_t) { return 0.0; }
double CPUAdapter::executeGCMStep(const WeatherGrid&, WeatherGrid&, scalar_t) { return 0.0; }
double CPUAdapter::calculateDiagnostics(WeatherGrid&) { return 0.0; }
}  // namespace weather_sim
