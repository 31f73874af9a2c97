#!/bin/bash
mkdir -p gpurun_out
python tools/pipeline_profile.py 64 256 > gpurun_out/r2m_pipeline_profile.txt 2>&1; head -120 gpurun_out/r2m_pipeline_profile.txt
