#!/bin/bash
python scripts/e2e_trace.py 1048576 3 2> gpurun_out/r2i_trace_1m_3.log; cat gpurun_out/r2i_trace_1m_3.log
python scripts/e2e_trace.py 2097152 3 2> gpurun_out/r2i_trace_2m_3.log; tail -8 gpurun_out/r2i_trace_2m_3.log
