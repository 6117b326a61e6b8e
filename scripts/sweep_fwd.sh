for cfg in "1 3" "2 3" "1 2" "2 2" "3 2"; do set -- $cfg
sed -i "s/#define GATHER_DEPTH .*/#define GATHER_DEPTH $1/; s/#define FWD_CTAS [0-9]*/#define FWD_CTAS $2/" neuralvolumetricreconstructionformedicalimages_b200/csrc/density_tc.cu
python -m neuralvolumetricreconstructionformedicalimages_b200.build >/dev/null 2>&1 || echo BUILD FAIL
echo "depth=$1 ctas=$2"; ATTRIB_FLAGS="0" bash scripts/attrib.sh
done
