#!/usr/bin/env python
"""Summarise an `ncu --set full` report of tc_contract_kernel into the small JSON bench.py reads
(profiles/r2_tc_contract_ncu.json): DRAM bytes per step (summed over the passes of one step),
tensor-pipe activity, issue activity, shared-memory wavefronts, local-memory requests.

    python scripts/dev/ncu_summary.py gpurun_out/<report>.ncu-rep profiles/r2_tc_contract_ncu.json [n_passes]
"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    npass = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if npass:
        data = data[:npass]
    col = {h: i for i, h in enumerate(hdr)}

    def val(name, r):
        v = float(r[col[name]])
        u = units[col[name]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3}.get(u, 1.0)

    t = [val("gpu__time_duration.sum", r) for r in data]
    w = [x / sum(t) for x in t]
    res = {
        "source": f"ncu --set full --clock-control none, {len(data)} launch(es) of tc_contract_kernel = one step "
                  f"(512^3, S=40, 6730 triangles); report {rep.split('/')[-1]}",
        "kernel_ms": t,
        "dram_bytes_per_step": sum(val("dram__bytes_read.sum", r) + val("dram__bytes_write.sum", r) for r in data),
        "dram_read_bytes": [val("dram__bytes_read.sum", r) for r in data],
        "dram_write_bytes": [val("dram__bytes_write.sum", r) for r in data],
        "tensor_pipe_active_pct": sum(wi * val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", r)
                                      for wi, r in zip(w, data)),
        "issue_active_pct": sum(wi * val("smsp__issue_active.avg.pct_of_peak_sustained_active", r) for wi, r in zip(w, data)),
        "smem_wavefront_pct": sum(wi * val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", r)
                                  for wi, r in zip(w, data)),
        "smem_ld_bank_conflicts": sum(val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", r) for r in data),
        "smem_ld_wavefronts": sum(val("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", r) for r in data),
        "local_load_requests": sum(val("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", r) for r in data),
        "local_store_requests": sum(val("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", r) for r in data),
        "registers_per_thread": val("launch__registers_per_thread", data[0]),
    }
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
