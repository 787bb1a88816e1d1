# r2as: knob re-check after the traversal trims (occupancy variants are the runtime-knob kernels; default = baked-knob kernels at 8 CTAs/SM)
bash tools/gpu_sweep.sh r2as "OHB_X=default;OHB_TRACE_OCC=7;OHB_TRACE_OCC=9;OHB_LANES=1;OHB_LANES=2" "helmet synthetic2m"
