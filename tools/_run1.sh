python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do
export SE3_AGG_PAIR=$v; echo "SE3_AGG_PAIR=$v"
python tools/layer_kernel_table.py 2>&1 | grep -E "seg_head|patch_dec0|dec2|enc0_block0|fpn0|enc1_block0|total"
done
