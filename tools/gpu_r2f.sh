#!/bin/bash
bash tools/gpu_verify.sh
bash tools/gpu_c4prof.sh
