"""Constants of the reference's token layout (madrigal/utils.py:28-37): ORDERED cell lines and non-TX modalities.

As in the reference, the environment variable NON_TX_MODALITIES ("str_kg_cv_bs", underscore-separated, read at import,
utils.py:30-34) sets the DEFAULT list of non-transcriptomic modalities; every drop-in class additionally takes the list
(or its length) as a constructor argument, so two encoders with different token layouts can live in one process."""
import os

CELL_LINES = ['a375', 'a549', 'asc', 'ha1e', 'hcc515', 'hec108', 'hela', 'hepg2', 'ht29', 'huvec', 'mcf7', 'npc',
              'pc3', 'thp1', 'vcap', 'yapc']
_env = os.getenv("NON_TX_MODALITIES")
NON_TX_MODALITIES = _env.split("_") if _env else ["str", "kg", "cv"]
NUM_NON_TX_MODALITIES = len(NON_TX_MODALITIES)
NUM_MODALITIES = NUM_NON_TX_MODALITIES + len(CELL_LINES)


def resolve_non_tx(non_tx_modalities=None):
    """Constructor-argument form of the knob: None -> the module default (environment / ['str','kg','cv']), an int ->
    that many non-TX tokens, a list -> the modality names (the first three must be str, kg, cv as in the reference's
    stacking order, models.py:772)."""
    if non_tx_modalities is None:
        return list(NON_TX_MODALITIES)
    if isinstance(non_tx_modalities, int):
        if non_tx_modalities < 3:
            raise ValueError("at least the str, kg and cv modalities are required")
        return ["str", "kg", "cv"] + [f"mod{i}" for i in range(3, non_tx_modalities)]
    mods = list(non_tx_modalities)
    if mods[:3] != ["str", "kg", "cv"]:
        raise ValueError("non_tx_modalities must start with ['str', 'kg', 'cv'] (models.py:772)")
    return mods
