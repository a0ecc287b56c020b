"""Special token ids (reference: utils/config.py:1-7)."""
PAD_TOKEN, UNK_TOKEN, BOS_TOKEN, EOS_TOKEN, SPC_TOKEN = '<pad>', '<unk>', '<s>', '</s>', '<spc>'
PAD, UNK, BOS, EOS, SPC = 0, 1, 2, 3, 4
