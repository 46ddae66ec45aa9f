import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d["extras"]
print("step_ms %.3f  C1 single %.3f ms batch %.0f reg/s | C3 single %.3f ms, 64 pairs %.0f reg/s pair_kernel %.3f ms | rot %.3f ms" % (d["ms_per_step"], e["C1_teapot_p2p_3d"]["single_call_ms"], e["C1_teapot_p2p_3d"]["batch_e2e_registrations_per_s"], e["C3_scan_to_submap"]["single_call_ms"], e["C3_scan_to_submap"]["e2e_registrations_per_s"], e["C3_scan_to_submap"]["pair_kernel_ms"], e["F1_rotation_search"]["single_call_ms"]))
