#!/bin/bash
# Build the A/B variants of the ICP iteration kernel for the first GPU call of the next round (run locally, the
# .so files travel with the snapshot: build_variants/ is git-ignored but not gpurun-ignored).
set -e
cd "$(dirname "$0")/.."
python -m cupoch_b200.build >/dev/null
cp cupoch_b200/lib/libcupoch_b200.so build_variants/base.so
for mb in 4 5 6 7 8; do tools/build_variant.sh mb$mb -DICP_MIN_BLOCKS=$mb; done
tools/build_variant.sh lowreg9 -DICP_LOWREG=1 -DICP_MIN_BLOCKS=9
tools/build_variant.sh lowreg8 -DICP_LOWREG=1 -DICP_MIN_BLOCKS=8
tools/build_variant.sh pdl -DCPHB_PDL=1
tools/build_variant.sh deep -DICP_DEEP_PIPE=1
tools/build_variant.sh deep_mb4 -DICP_DEEP_PIPE=1 -DICP_MIN_BLOCKS=4
tools/build_variant.sh faststart -DICP_FAST_START=1
tools/build_variant.sh deep_fast_mb4 -DICP_DEEP_PIPE=1 -DICP_FAST_START=1 -DICP_MIN_BLOCKS=4
tools/build_variant.sh deep_fast_pdl_mb4 -DICP_DEEP_PIPE=1 -DICP_FAST_START=1 -DCPHB_PDL=1 -DICP_MIN_BLOCKS=4
tools/build_variant.sh deep_pdl -DICP_DEEP_PIPE=1 -DCPHB_PDL=1
tools/build_variant.sh pdl_mb6 -DCPHB_PDL=1 -DICP_MIN_BLOCKS=6
for sb in 6 7 8; do tools/build_variant.sh dual_s$sb -DICP_DUAL=1 -DICP_DEEP_PIPE=1 -DICP_FAST_START=1 -DICP_MIN_BLOCKS=4 -DICP_MIN_BLOCKS_SEARCH=$sb; done
tools/build_variant.sh lane_dual_s6_b3 -DICP_LANE_ACC=1 -DICP_DUAL=1 -DICP_DEEP_PIPE=1 -DICP_FAST_START=1 -DICP_MIN_BLOCKS=3 -DICP_MIN_BLOCKS_SEARCH=6
tools/build_variant.sh lane_dual_s6_b4 -DICP_LANE_ACC=1 -DICP_DUAL=1 -DICP_DEEP_PIPE=1 -DICP_FAST_START=1 -DICP_MIN_BLOCKS=4 -DICP_MIN_BLOCKS_SEARCH=6
ls -la build_variants
