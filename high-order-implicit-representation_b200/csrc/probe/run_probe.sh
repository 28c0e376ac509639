P=high-order-implicit-representation_b200/csrc/probe/umma_probe
run() { timeout 20 $P "$@" 2>&1 | grep -E "^M=|error" | sed -e 's/max_err.*| /| /'; rc=${PIPESTATUS[0]}; [ $rc -ne 0 ] && echo "  (exit $rc for: $*)"; }
echo "== uniform-warp issue (swapA=8), SS K-major none, nacc = 1/4"
for n in 16 40 64 128 208 256; do run 128 $n 64 0 0 8 1 0 200 0 0; done
for n in 16 40 64; do run 128 $n 64 0 0 8 4 0 200 0 0; done
echo "== uniform-warp issue, TS"
for n in 16 40 64 208; do run 128 $n 64 0 0 8 1 0 200 9 0; done
for n in 16 40; do run 128 $n 64 0 0 8 4 0 200 9 0; done
echo "== M=64"
for n in 40 208; do run 64 $n 64 0 0 8 1 0 200 0 0; done
