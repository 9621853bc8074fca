import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        continue
    d=json.loads(line)
    k=d["kernel_ms"]
    u=max(1,d["roofline"]["units"])
    print("value %.0f scen/s  ms/step %.1f  conv %d  it/scen %.2f  factor %.0f solve %.0f jac %.0f mis %.0f other %.0f | front frac %.3f step frac %.3f iters/s %.0f | factor us/unit %.2f solve us/unit %.2f units %d" % (d["value"],d["ms_per_step"],d["converged"],d["iterations_per_scenario"],k["factor_ms"],k["solve_ms"],k["jacobian_ms"],k["mismatch_ms"],k["other_ms"],d["roofline"]["frac"],d["roofline_step"]["frac"],d["roofline_step"]["scenario_iterations_per_s"],1e3*k["factor_ms"]/u,1e3*k["solve_ms"]/u,u))
