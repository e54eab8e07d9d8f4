#!/bin/bash
# build the library and verify it loads; prints BUILD OK / BUILD FAILED
cd "$(dirname "$0")/.."
if python -m meme_challenge_b200.build > /tmp/b200u_build.log 2>&1 && python -c "import ctypes; ctypes.CDLL('meme_challenge_b200/libb200u.so')"; then echo BUILD OK; else grep -B2 -A6 "error" /tmp/b200u_build.log | head -40; echo BUILD FAILED; fi
