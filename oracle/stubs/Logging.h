// Stand-in for the reference's spdlog-based include/Logging.h (spdlog is not in this image): the oracle/_ref build
// compiles /root/reference/src/DescriptorPool.cc in place and only needs its two log macros to vanish.
// TEST INFRASTRUCTURE (see oracle/__init__.py).
#pragma once
// Same include guard as the real header: force-included first (-include), it turns the real Logging.h - which sits
// next to Profiling.h and would win the quote-include lookup - into a no-op.
#define SUPERSLAM_LOGGING_H
#define SLOG_TRACE(...) ((void)0)
#define SLOG_CRITICAL(...) ((void)0)
#define SLOG_ERROR(...) ((void)0)
#define SLOG_WARN(...) ((void)0)
#define SLOG_INFO(...) ((void)0)
#define SLOG_DEBUG(...) ((void)0)
