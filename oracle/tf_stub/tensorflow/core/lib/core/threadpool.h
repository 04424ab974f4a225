// TEST INFRASTRUCTURE ONLY -- stand-in for tensorflow/core/lib/core/threadpool.h: the shim runs shards serially.
#pragma once
namespace tensorflow {
namespace thread {
class ThreadPool {};
}  // namespace thread
}  // namespace tensorflow
