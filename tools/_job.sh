set +e
tools/profile.sh kernel lines_trace_hall "k_trace" 12 1
tools/profile.sh kernel lines_primary_hall "k_mesh_primary" 3 1
ls -la gpurun_out/*.ncu-rep
