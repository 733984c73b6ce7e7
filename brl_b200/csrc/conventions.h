// conventions.h -- pgx 1.4.0 behaviours that NOTHING in the reference tree pins
// ("parity unpinned").  Kept in one switchable place so they
// can be flipped if a pgx install ever shows otherwise.  None of them is observable
// through brl's consumers: auto_reset / duplicate_init overwrite a terminal state
// (src/utils.py:45-55, src/duplicate.py:151-155) and the evaluation loops ignore it.
#pragma once

// pgx core sets legal_action_mask to all-True on a terminated state.
#define BRL_CONV_TERMINAL_MASK_ALL_TRUE 1
// current_player is NOT advanced by the call that ends the auction.
#define BRL_CONV_TERMINAL_ADVANCES_PLAYER 0
