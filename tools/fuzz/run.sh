#!/bin/bash
# Build the parsers with ASan + UBSan and fuzz them, then the watchdog under TSan (CPU only, ~1 min): tools/fuzz/run.sh [iterations]
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
W=$(mktemp -d)
cd $ROOT/birda_b200/csrc
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -c wav.cpp -o $W/wav.o
nvcc -O1 -g -std=c++17 --expt-relaxed-constexpr -diag-suppress 20011,20014 -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer -c k6_flac.cu -o $W/k6.o
cd $ROOT/tools/fuzz
nvcc -Wno-deprecated-gpu-targets -o $W/drv parsers_asan.cpp stubs.cpp $W/wav.o $W/k6.o -Xcompiler -fsanitize=address,-fsanitize=undefined -Xlinker -lasan,-lubsan
cd $ROOT
python - "$W" <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from birda_b200.synth import synth_pcm, write_wav
import flac_enc
w = sys.argv[1]
write_wav(w + "/a.wav", synth_pcm(1, 0.05, 16000, 2), 16000, 2)
pc = synth_pcm(3, 0.5, 16000, 2).reshape(-1, 2)
open(w + "/a.flac", "wb").write(flac_enc.encode(pc, 16000, 16, style=dict(kinds=["lpc"], stereo="mid_side", part_order=2)))
PY
ASAN_OPTIONS=detect_leaks=0 $W/drv $W/a.wav $W/a.flac ${1:-20000}
cd $ROOT/birda_b200/csrc
g++ -O1 -g -std=c++17 -fsanitize=thread -fPIC -pthread -c watchdog.cpp -o $W/watchdog.o
cd $ROOT/tools/fuzz
g++ -O1 -g -std=c++17 -fsanitize=thread -pthread -o $W/wd watchdog_tsan.cpp stubs.cpp $W/watchdog.o
$W/wd
cd $ROOT/birda_b200/csrc
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -c mask.cpp -o $W/mask.o
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fPIC -fno-fast-math -ffp-contract=off -c rules.cpp -o $W/rules.o
cd $ROOT/tools/fuzz
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -o $W/mk mask_asan.cpp stubs.cpp $W/mask.o $W/rules.o
$W/mk
rm -rf $W
