#!/bin/bash
bash tools/gpu_job_prof.sh
bash tools/gpu_job_san.sh
