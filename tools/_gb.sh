for w in 12; do for sp in 1 2; do SCHEMANET_GRAPH_WARPS=$w SCHEMANET_GRAPH_SPLIT=$sp timeout 60 python tools/graph_bench.py 256 1024; done; done
for w in 12; do SCHEMANET_GRAPH_WARPS=$w timeout 60 python tools/graph_bench.py 1024 8000; SCHEMANET_GRAPH_WARPS=$w timeout 60 python tools/graph_bench.py 64 128; SCHEMANET_GRAPH_WARPS=$w timeout 60 python tools/graph_bench.py 512 1024; done
SCHEMANET_GRAPH_WARPS=8 timeout 60 python tools/graph_bench.py 512 1024
