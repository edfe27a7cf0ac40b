# round-2 profile captures (run on the GPU box): launch list + ncu --set full of the kernels DESIGN.md quotes
set +e
tools/profile.sh launches r2_hall
tools/profile.sh kernel r2_trace_hall "k_trace" 12 2
tools/profile.sh kernel r2_primary_hall "k_mesh_primary" 3 1
tools/profile.sh kernel r2_shade_hall "k_shade" 10 3
tools/profile.sh kernel r2_trace_1m "k_trace" 6 1 --workload scene_1m_1080p
tools/profile.sh kernel r2_primary_1m "k_mesh_primary" 3 1 --workload scene_1m_1080p
tools/profile.sh kernel r2_trace_10m "k_trace" 24 3 --workload scene_10m_4k --steps 1
tools/profile.sh kernel r2_build_1m "k_ploc_loop|k_collapse_loop|k_rs_scatter|k_rs_hist|k_emit_nodes" 0 12 --workload scene_1m_1080p
tools/profile.sh kernel r2_spheres "k_spheres" 4 2 --workload spheres_960x540
tools/profile.sh kernel r2_path_hall "k_path" 3 1 --opt path_kernel=1
ls -la gpurun_out/*.txt gpurun_out/launches_r2_hall.csv | tail -20
rm -f gpurun_out/*.ncu-rep
