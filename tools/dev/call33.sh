#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  bash tools/measure_round.sh
} 2>&1 | tee gpurun_out/call33.log
