"""Constants of the reference's token layout (madrigal/utils.py:28-37): ORDERED cell lines and non-TX modalities."""
CELL_LINES = ['a375', 'a549', 'asc', 'ha1e', 'hcc515', 'hec108', 'hela', 'hepg2', 'ht29', 'huvec', 'mcf7', 'npc',
              'pc3', 'thp1', 'vcap', 'yapc']
NON_TX_MODALITIES = ["str", "kg", "cv"]
NUM_NON_TX_MODALITIES = len(NON_TX_MODALITIES)
NUM_MODALITIES = NUM_NON_TX_MODALITIES + len(CELL_LINES)
