for F in 148 296 592 1024 2048; do ILCC2_LIB=build_variants/lib_timing.so python tools/fit_timing.py $F 2>&1 | head -3; done
