// TEST INFRASTRUCTURE: src/audio.h names ZSTD_CStream in a member declaration only.
#pragma once
typedef struct ZSTD_CCtx_s ZSTD_CStream;
