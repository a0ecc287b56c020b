set -u
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/one_step_graph.py > $O/one_step_graph.log 2>&1 || { echo "one_step_graph failed"; tail -5 $O/one_step_graph.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02k_graph_launches.csv python scripts/one_step_graph.py > $O/ncu_list_graph.log 2>&1
echo "rows: $(grep -c '^"' $O/r02k_graph_launches.csv)"
python scripts/summarize_graph_launches.py $O/r02k_graph_launches.csv > $O/r02k_graph_launches_summary.txt; head -12 $O/r02k_graph_launches_summary.txt
