"""Memory-system detail of ONE kernel from an `ncu --set full` report (read on the CPU box):
    python profiles/ncu_kernel_detail.py gpurun_out/prof.ncu-rep gb_bwd_kernel > profiles/ncu_rNN_gb_bwd_detail.txt
L1TEX requests / sectors by space and operation (global ld / st / red, local = register spills), L2 requests, unit
utilisations, issue statistics - the numbers DESIGN.md quotes when it says what does NOT bound a kernel."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st_lookup_miss.sum", "SM_B.TriageCompute.l1tex__t_sectors.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_requests_srcunit_tex_op_write.sum",
    "lts__t_sectors.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_ltcfabric.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main(path, kernel):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if kernel in r[ki]:
            print("# %s" % r[ki].replace("<unnamed>::", "")[:160])
            print("# from %s (ncu --set full, cold caches, serialised: shapes and ratios, not bench times)" % path)
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    print("%-88s %-10s %s" % (m, units[i], r[i]))
            print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
