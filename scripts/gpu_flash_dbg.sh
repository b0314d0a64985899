# knock-out timing of the flash attention backward (OG_FLASH_DBG bits: 1 no MUFU, 2 no P/dS stores, 4 no gradient MMAs, 8 no S/dP MMAs)
for d in ${DBGS:-0 4 12 15}; do echo "DBG=$d $(OG_FLASH_DBG=$d timeout 120 python scripts/ncu_attention.py | tail -1 | cut -c1-75)"; done
