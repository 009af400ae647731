for v in base e1 e2 e4 e8 e16 e31; do echo "== $v"; BALF_B200_LIB=$PWD/build/variants/$v.so python scripts/variant_ab.py 64 0x01 2>&1 | grep -E "^mask|c32"; done
