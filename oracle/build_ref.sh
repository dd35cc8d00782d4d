#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *real* reference CPU path so the oracle
# restatement (oracle/ws_oracle.c) and the golden vectors can be pinned against it.
#
# The reference weather-sim does not compile as shipped, so the
# recipe patches a THROW-AWAY copy under $TMPDIR (never inside this repo), compiles
# it from there, and leaves only binaries in oracle/_ref/:
#   oracle/_ref/libws_ref.so            reference classes behind oracle/ref_shim.cpp (ctypes)
#   oracle/_ref/pyweather_sim.*.so      the reference's own pybind11 module (surface checks)
#   oracle/_ref/weather_simulation.pyc.bin  the reference's own Python wrapper, byte-compiled unmodified (the
#                                        ".bin" keeps snapshot tools that drop *.pyc from losing it)
#
# The six patches are mechanical compile fixes; none
# touches arithmetic. -ffp-contract=off is mandatory (GCC defaults to "fast").
#
# Needs /root/reference (dev container only). On the GPU box the prebuilt files travel
# with the snapshot; this script is then a no-op.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${WS_REFERENCE_ROOT:-/root/reference}/src/weather-sim/cpp"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
mkdir -p "$OUT"

if [ ! -d "$REF" ]; then
    echo "build_ref: $REF not present; keeping prebuilt oracle/_ref (if any)" >&2
    exit 0
fi

EXT_SUFFIX="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
if [ "${1:-}" != "--force" ] && [ -f "$OUT/libws_ref.so" ] && [ -f "$OUT/pyweather_sim$EXT_SUFFIX" ] && [ -f "$OUT/weather_simulation.pyc.bin" ] \
   && [ "$OUT/libws_ref.so" -nt "$HERE/ref_shim.cpp" ] && [ "$OUT/libws_ref.so" -nt "$HERE/build_ref.sh" ]; then
    exit 0
fi

WORK="$(mktemp -d "${TMPDIR:-/tmp}/ws_ref_build.XXXXXX")"
trap 'rm -rf "$WORK"' EXIT
mkdir -p "$WORK/include/weather_sim" "$WORK/src"
cp "$REF"/include/weather_sim/{weather_sim.hpp,initial_conditions.hpp,gpu_adaptability.hpp,output_manager.hpp} "$WORK/include/weather_sim/"
cp "$REF"/src/{weather_grid.cpp,weather_simulation.cpp,initial_conditions.cpp,gpu_adaptability.cpp,python_bindings.cpp} "$WORK/src/"

# P1: WeatherGrid declares `index_t height_` and `ScalarField2D height_` (weather_sim.hpp:397 vs :404)
sed -i -e '397s/index_t height_;/index_t grid_height_;/' \
       -e '285s/return height_;/return grid_height_;/' "$WORK/include/weather_sim/weather_sim.hpp"
# P2: follow the rename in weather_grid.cpp (first height_( initialiser of each ctor, loop bounds, swap)
sed -i -e '17s/height_(height)/grid_height_(height)/' \
       -e '38s/height_(config.grid_height)/grid_height_(config.grid_height)/' \
       -e '87s/height_/grid_height_/' -e '93s/height_ - 1/grid_height_ - 1/' \
       -e '105s/height_/grid_height_/' -e '111s/height_ - 1/grid_height_ - 1/' \
       -e '125s/height_ != other.height_/grid_height_ != other.grid_height_/' "$WORK/src/weather_grid.cpp"
# P3: missing standard includes in initial_conditions.hpp
sed -i -e '13a #include <map>\n#include <vector>\n#include <type_traits>' "$WORK/include/weather_sim/initial_conditions.hpp"
# P4: RK2 uses T/p references that are out of scope at weather_simulation.cpp:315-317
sed -i -e '313a\        auto\& current_temp = current_grid_->getTemperatureField();\n        auto\& tendency_temp = tendency_grid_->getTemperatureField();\n        auto\& current_pressure = current_grid_->getPressureField();\n        auto\& tendency_pressure = tendency_grid_->getPressureField();' "$WORK/src/weather_simulation.cpp"
# P6: ambiguous overload in the pybind11 .def (python_bindings.cpp:351)
sed -i -e '351s/&WeatherSimulation::getCurrentGrid/py::overload_cast<>(\&WeatherSimulation::getCurrentGrid)/' "$WORK/src/python_bindings.cpp"
# P5 is oracle/ref_stub.cpp (link stub for the declared-but-undefined CPUAdapter).

CXXFLAGS="-std=c++17 -O2 -ffp-contract=off -fopenmp -fPIC -w"
COMMON="$WORK/src/weather_grid.cpp $WORK/src/weather_simulation.cpp $WORK/src/initial_conditions.cpp $WORK/src/gpu_adaptability.cpp $HERE/ref_stub.cpp"

g++ $CXXFLAGS -shared -I"$WORK/include" $HERE/ref_shim.cpp $COMMON -o "$OUT/libws_ref.so" &
PID1=$!
g++ $CXXFLAGS -shared -I"$WORK/include" -I"$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')" \
    -I"$($PY -c 'import pybind11;print(pybind11.get_include())')" \
    "$WORK/src/python_bindings.cpp" $COMMON -o "$OUT/pyweather_sim$EXT_SUFFIX" &
PID2=$!
wait $PID1
wait $PID2
# The reference's own Python wrapper (python/weather_simulation.py), UNMODIFIED, byte-compiled: a binary artefact
# like the two libraries above (no reference source enters the repo). tests/test_reference_wrapper.py puts it next
# to the B200 pyweather_sim module and next to the reference's, and runs the same user script on both.
"$PY" - "$REF/../python/weather_simulation.py" "$OUT/weather_simulation.pyc.bin" <<'PYEOF'
import py_compile, sys
py_compile.compile(sys.argv[1], cfile=sys.argv[2], dfile="reference:src/weather-sim/python/weather_simulation.py", doraise=True)
PYEOF
echo "build_ref: wrote $OUT/libws_ref.so, $OUT/pyweather_sim$EXT_SUFFIX and $OUT/weather_simulation.pyc.bin"
