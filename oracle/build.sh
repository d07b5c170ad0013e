#!/usr/bin/env bash
# Builds the CPU oracle (test infrastructure, see t5_oracle.cpp header).
# -ffp-contract=off: the specification forbids fused multiply-add.
# oracle/_ref (the reference compiled from its own sources) is NOT built: the
# reference is pure Haskell and this image has no ghc/cabal/stack.
set -euo pipefail
cd "$(dirname "$0")"
g++ -std=c++17 -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -pthread \
    -o libt5oracle.so t5_oracle.cpp
echo "built oracle/libt5oracle.so"
